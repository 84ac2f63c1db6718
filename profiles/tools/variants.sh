T="tests/test_gpu_parity.py::test_gradients_match_reference_autograd"
for v in "B2S_FWD_EX2=1" "B2S_FWD_EX2=1 B2S_BWD_EX2=1" "B2S_FWD_EX2=1 B2S_BWD_MMASYNC=1" "B2S_FWD_MMASYNC=1 B2S_BWD_EX2=1"; do
  echo "=== $v"; env $v python -m pytest "$T" -q -k "depth-r1_many_small or depth-r1_sh4_bg" 2>&1 | grep -E "^E   +assert [0-9]|passed|failed" | head -6
done
