// Every environment variable libvaura_b200.so reads, in one place.
//
// None of them is part of the interface a caller needs: the defaults are the measured-best configuration and are what the
// tests, smoke() and bench.py run.  They exist for A/B measurements (profiles/scripts/*.sh set them) and for the profiling
// scripts under profiles/.  Each extern "C" entry point calls refresh_knobs() first, so a change of the environment takes
// effect at the next API call of the calling thread; the code below the entry points only reads knobs().
#pragma once

namespace vaura {

struct Knobs {
  // ---- sampler dispatch (cabi.cu) ----
  int deterministic;       // VAURA_DETERMINISTIC=1      bf16 step kernel: split-K slices added in a fixed order (11 % slower)
  int prefill_tc;          // VAURA_PREFILL_TC=0         multi-position fp32-activation passes stay on the GEMV kernels
  int fused_step;          // VAURA_FUSED_STEP=0         bf16 decode step as a graph of ~170 kernels instead of one kernel
  int fused_io;            // VAURA_FUSED_IO=0           embedding / sampling as separate launches around the fused step
  int fused2;              // VAURA_FUSED2=1             second fused design (decode_fused2.cu; parity-green, slower)
  int fused2_flags;        // VAURA_FUSED2_FLAGS=n       bit 0: hold the weight ring back until the barrier opens
  int fused2_nocoop;       // VAURA_FUSED2_NOCOOP=1      cluster attribute only (Nsight Compute rejects cooperative clusters)
  int bf16_step_first;     // VAURA_BF16_STEP_FIRST=0    separate first pass over position 0 on the bf16 path
  int no_persistent;       // VAURA_NO_PERSISTENT=1      rows <= 4: graph-replayed multi-kernel step
  int no_cluster;          // VAURA_NO_CLUSTER=1         rows <= 2: decode_step_persistent instead of decode_step_cluster
  int cluster_nocoop;      // VAURA_CLUSTER_NOCOOP=1     as fused2_nocoop, for decode_step_cluster
  int cluster_ring;        // VAURA_CLUSTER_RING=n       cluster kernel: extra L2 prefetch distance (measured slower)
  int cluster_l2_ahead;    // VAURA_CLUSTER_L2_AHEAD=n   cluster kernel: pacing of the L2 prefetch in cycles (-1 = built-in)
  int cluster_tail_units;  // VAURA_CLUSTER_TAIL_UNITS=n cluster kernel: units of the next step prefetched during the tail
  int persist_prefetch;    // VAURA_PERSIST_PREFETCH=n   persistent kernel: groups of L2 prefetch ahead (measured: no gain)
  int any_page;            // VAURA_ANY_PAGE=1           accept any power-of-two K/V page >= 16 on the bf16 path (experiment)
  int phase_timing;        // VAURA_PERSIST_TIMING=1     step kernels write phase timestamps into the workspace (profiles/*_timing.py)
  int timing_cta;          // VAURA_TIMING_CTA=n         ... of this CTA
  // ---- unfused bf16 pass (first pass with a prompt, rows > 128) ----
  int pdl_mode;            // VAURA_PDL_MODE=bits        programmatic dependent launch experiments (off: no gain)
  int no_splitk;           // VAURA_NO_SPLITK=1
  int wo_bn, wo_ksplit;    // VAURA_WO_BN / VAURA_WO_KSPLIT   tile width and K split of the wo GEMM (64, 6)
  int w2_bn, w2_ksplit;    // VAURA_W2_BN / VAURA_W2_KSPLIT   ... of the w2 GEMM (64, 6)
  // ---- fused bf16 step kernel (gemm_tcgen05.cu) ----
  int fused_l2_prefetch;   // VAURA_FUSED_L2_PREFETCH=0  no L2 prefetch of the next phase's weight tile
  int fused_tm128;         // VAURA_FUSED_TM128=1        UMMA M = 128 also for <= 64 rows
  // ---- tensor-core prefill ----
  int prefill_bf16;        // VAURA_PREFILL_BF16=0       a sampling call keeps the three-term (fp32-equivalent) prompt prefill
  int prefill_ck;          // VAURA_PREFILL_CK=0         no K split inside clusters
  int prefill_attn_qw;     // VAURA_PREFILL_ATTN_QW=1|2|4 queries per warp of the fp32 prefill attention (0 = by grid size)
  int prefill_bn256_from;  // VAURA_PREFILL_BN256_FROM=n 256-wide tiles from this N (8192)
  // ---- codec ----
  int codec_simt;          // VAURA_CODEC_SIMT=1         every convolution on the CUDA-core kernel (read when a codec is created)
  int codec_fused_ru;      // VAURA_CODEC_FUSED_RU=0     residual units as two launches
  int codec_ru_skew;       // VAURA_CODEC_RU_SKEW=0      encoder's 64- / 128-channel fused residual units without the one-tile skew
  int conv_persistent;     // VAURA_CONV_PERSISTENT=0    one CTA per tile instead of the persistent tile loop
  int conv_ksub;           // VAURA_CONV_KSUB=0
  int conv_occ2;           // VAURA_CONV_OCC2=1          two CTAs per SM for the 96-channel layers (measured neutral)
  // ---- Segment-AVCLIP tower ----
  int avclip_simt_attn;    // VAURA_AVCLIP_SIMT_ATTN=1   SIMT space attention for every shape
  int avclip_m_fastest;    // VAURA_AVCLIP_M_FASTEST=1   the codec's tile order
  int avclip_2cta;         // VAURA_AVCLIP_2CTA=0        one CTA per tile instead of CTA pairs
  int avclip_ew8;          // VAURA_AVCLIP_EW8=1         eight epilogue warps
};

const Knobs& knobs();   // the calling thread's copy (defaults until refresh_knobs() has run on this thread)
void refresh_knobs();   // re-read the environment

}  // namespace vaura
