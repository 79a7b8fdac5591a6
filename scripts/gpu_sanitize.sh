# compute-sanitizer memcheck over the Viterbi paths (all three mappings) and the FIC self-configuration test
cd $GRAFT_REPO_ROOT
TAG=${1:-san}
(timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_viterbi_matches_oracles" "tests/test_channel_gpu.py::test_frame_decode_fic_msc_dabplus" tests/test_autoconfig_gpu.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${TAG}_memcheck.log 2>&1
(DABGPU_VIT_LANE_CTAS_PER_SM=4 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_viterbi_matches_oracles" -m gpu -x -q 2>&1 | tail -12) > gpurun_out/${TAG}_memcheck_wide.log 2>&1
cat gpurun_out/${TAG}_memcheck.log gpurun_out/${TAG}_memcheck_wide.log
