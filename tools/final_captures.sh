set -x

ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_ncu_verify.csv python tools/profile_verify.py llama-3.1-8b-q4km 2048 2 > gpurun_out/r2k_ncu_verify.log 2>&1
python tools/ncu_summary.py gpurun_out/r2k_ncu_verify.csv --from-last embed_kernel --grid 2048 --meta shape=llama-3.1-8b-q4km tokens=2048 "command=ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none python tools/profile_verify.py llama-3.1-8b-q4km 2048 2 (last verify)" > gpurun_out/r2k_ncu_verify_summary.json
ncu --set full --clock-control none --import-source on -k regex:prefill_attn_tc -s 40 -c 1 -o gpurun_out/r2k_attn -f python tools/profile_verify.py llama-3.1-8b-q4km 2048 1 > gpurun_out/r2k_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefill_gemm_kernel -s 200 -c 4 -o gpurun_out/r2k_gemm -f python tools/profile_verify.py llama-3.1-8b-q4km 2048 1 > gpurun_out/r2k_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 8 -c 1 -o gpurun_out/r2k_mega -f python tools/profile_step.py llama-3.1-8b-q4km 12 512 > gpurun_out/r2k_mega.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2k_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --verify 0 > gpurun_out/r2k_ncu_bench.log 2>&1

ls -la gpurun_out/r2k_*
