# compute-sanitizer over the Viterbi paths (all three mappings, both lane launch shapes, the short and the general branch-error
# form), the time de-interleave / de-puncture pass, the DAB+ stage, the packet-mode FEC kernel, the FIC self-configuration test and
# the OFDM kernels: memcheck (out-of-bounds / misaligned accesses) and racecheck (shared-memory hazards of k_ofdm_demod2,
# k_ofdm_ctl, k_vit_prep, k_viterbi, k_dabplus)
cd $GRAFT_REPO_ROOT
TAG=${1:-san}
(timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_viterbi_matches_oracles" "tests/test_channel_gpu.py::test_viterbi_lane_plan_edge_cases" "tests/test_channel_gpu.py::test_frame_decode_fic_msc_dabplus" "tests/test_channel_gpu.py::test_dabplus_with_byte_errors" tests/test_autoconfig_gpu.py tests/test_packet_fec_gpu.py "tests/test_ofdm_gpu.py::test_ofdm_multi_stream_ragged_and_post_viterbi_identical" -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${TAG}_memcheck.log 2>&1
(DABGPU_VIT_LANE_CTAS_PER_SM=4 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_viterbi_matches_oracles" -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${TAG}_memcheck_wide.log 2>&1
(timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest "tests/test_channel_gpu.py::test_frame_decode_fic_msc_dabplus" "tests/test_ofdm_gpu.py::test_ofdm_multi_stream_ragged_and_post_viterbi_identical" -m gpu -x -q -k "vit-lanes or ragged" 2>&1 | tail -8) > gpurun_out/${TAG}_racecheck.log 2>&1
cat gpurun_out/${TAG}_memcheck.log gpurun_out/${TAG}_memcheck_wide.log gpurun_out/${TAG}_racecheck.log
