for k in k_big_max k_big_sum k_big_norm_est k_big_search_gather; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s 4 -c 1 -f -o gpurun_out/prof_$k python tools/profile_c5.py 1048576 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log | cut -c1-100
done
