// knobs.h: the environment variables of libvaura_b200.so, read into a per-thread struct
#include "knobs.h"

#include <cstdlib>

namespace vaura {
namespace {
int env_flag(const char* name, int dflt) {  // "1" / "0"; anything else (or unset) = default
  const char* e = getenv(name);
  if (!e || (e[0] != '0' && e[0] != '1')) return dflt;
  return e[0] == '1';
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
Knobs read_env() {
  Knobs k{};
  k.deterministic = env_flag("VAURA_DETERMINISTIC", 0);
  k.prefill_tc = env_flag("VAURA_PREFILL_TC", 1);
  k.fused_step = env_flag("VAURA_FUSED_STEP", 1);
  k.fused_io = env_flag("VAURA_FUSED_IO", 1);
  k.fused2 = env_flag("VAURA_FUSED2", 0);
  k.fused2_flags = env_int("VAURA_FUSED2_FLAGS", 0);
  k.fused2_nocoop = env_flag("VAURA_FUSED2_NOCOOP", 0);
  k.bf16_step_first = env_flag("VAURA_BF16_STEP_FIRST", 1);
  k.no_persistent = env_flag("VAURA_NO_PERSISTENT", 0);
  k.no_cluster = env_flag("VAURA_NO_CLUSTER", 0);
  k.cluster_nocoop = env_flag("VAURA_CLUSTER_NOCOOP", 0);
  k.cluster_ring = env_int("VAURA_CLUSTER_RING", 0);
  k.cluster_l2_ahead = env_int("VAURA_CLUSTER_L2_AHEAD", -1);
  k.cluster_tail_units = env_int("VAURA_CLUSTER_TAIL_UNITS", 0);
  k.persist_prefetch = env_int("VAURA_PERSIST_PREFETCH", 0);
  if (k.persist_prefetch < 0) k.persist_prefetch = 0;
  k.any_page = env_flag("VAURA_ANY_PAGE", 0);
  k.phase_timing = env_flag("VAURA_PERSIST_TIMING", 0);
  k.timing_cta = env_int("VAURA_TIMING_CTA", 0);
  k.pdl_mode = env_int("VAURA_PDL_MODE", 0);
  k.no_splitk = env_flag("VAURA_NO_SPLITK", 0);
  k.wo_bn = env_int("VAURA_WO_BN", 64);
  k.wo_ksplit = env_int("VAURA_WO_KSPLIT", 6);
  k.w2_bn = env_int("VAURA_W2_BN", 64);
  k.w2_ksplit = env_int("VAURA_W2_KSPLIT", 6);  // 24 x 6 = 144 CTAs: one wave (8 -> 192 CTAs was 6 % slower)
  k.fused_l2_prefetch = env_int("VAURA_FUSED_L2_PREFETCH", 1);
  k.fused_tm128 = env_flag("VAURA_FUSED_TM128", 0);
  k.prefill_bf16 = env_flag("VAURA_PREFILL_BF16", 1);
  k.prefill_ck = env_flag("VAURA_PREFILL_CK", 1);
  k.prefill_attn_qw = env_int("VAURA_PREFILL_ATTN_QW", 0);
  k.prefill_bn256_from = env_int("VAURA_PREFILL_BN256_FROM", 8192);
  k.codec_simt = env_flag("VAURA_CODEC_SIMT", 0);
  k.codec_fused_ru = env_flag("VAURA_CODEC_FUSED_RU", 1);
  k.codec_ru_skew = env_flag("VAURA_CODEC_RU_SKEW", 1);
  k.conv_persistent = env_flag("VAURA_CONV_PERSISTENT", 1);
  k.conv_ksub = env_flag("VAURA_CONV_KSUB", 1);
  k.conv_occ2 = env_flag("VAURA_CONV_OCC2", 0);
  k.avclip_simt_attn = env_flag("VAURA_AVCLIP_SIMT_ATTN", 0);
  k.avclip_m_fastest = env_flag("VAURA_AVCLIP_M_FASTEST", 0);
  k.avclip_2cta = env_flag("VAURA_AVCLIP_2CTA", 1);
  k.avclip_ew8 = env_flag("VAURA_AVCLIP_EW8", 0);
  return k;
}
thread_local Knobs t_knobs = read_env();
}  // namespace

const Knobs& knobs() { return t_knobs; }
void refresh_knobs() { t_knobs = read_env(); }

}  // namespace vaura
