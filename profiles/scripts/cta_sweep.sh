for c in 0 1 37 74 100 143 147; do
  echo "== timing cta $c"
  VAURA_TIMING_CTA=$c python profiles/fused_timing.py 64 127 2>&1 | grep -E "work" | head -7
done
