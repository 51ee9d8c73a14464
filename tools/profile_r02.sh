#!/bin/bash
# Round-2 ncu captures (one GPU, after the plain commands exited 0): the final K4, and the kernels of a streamed (multi-round) build.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
$CMD > gpurun_out/plain_k4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_v2.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:align_fast_ -s 3 -c 1 -f -o gpurun_out/r02_align_fast_v2 $CMD > gpurun_out/ncu_k4.log 2>&1
python tools/ncu_lines.py gpurun_out/r02_align_fast_v2.ncu-rep 70 > gpurun_out/r02_align_fast_ncu_v2_lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/r02_align_fast_v2.ncu-rep > gpurun_out/r02_align_fast_ncu_v2.json 2>gpurun_out/ncu_summary.err
# streamed build: 400 genomes x 2 Mb on one GPU = 2 rounds
cat > gpurun_out/streamed_build.py <<'PY'
import sys, os, json, time
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "bioinformatics-project-for-shotgun-metagenomics-pseudo-alignment-shotgun-_b200")]
import numpy as np, torch, bench, multi_gpu, _native as nat
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
G, GL = 400, 2_000_000
bases = bench.device_genomes(torch, dev, G, GL, seed=5000)
goff = (np.arange(G + 1, dtype=np.uint64) * np.uint64(GL))
torch.cuda.synchronize()
for it in range(2):
    t0 = time.perf_counter()
    dix = multi_gpu.build_partitioned(None, bases.data_ptr(), goff, 31, device=0, table_only=True, n_rounds=2, g_range=(0, G))
    torch.cuda.synchronize()
    print(json.dumps({"seconds": time.perf_counter() - t0, "phases": dix.timings, "keys": int(dix.replica.info().n_keys), "table_bytes": int(dix.replica.info().table_bytes)}))
    dix.close()
PY
python gpurun_out/streamed_build.py > gpurun_out/streamed_build.log 2>&1 || { tail -5 gpurun_out/streamed_build.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'owner_scatter|table_insert|encode_windows|owner_count' -s 0 -c 4 -f -o gpurun_out/r02_streamed_build python gpurun_out/streamed_build.py > gpurun_out/ncu_build.log 2>&1
for k in owner_scatter table_insert encode_windows; do
  ncu -i gpurun_out/r02_streamed_build.ncu-rep --page raw --csv -k regex:$k 2>/dev/null | python -c "
import csv, sys, json
rows = list(csv.reader(sys.stdin))
if len(rows) < 3: sys.exit(0)
h = rows[0]; r = rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'lts__t_bytes_srcunit_tex_op_write.sum']
print(json.dumps({k: r[h.index(k)] for k in want if k in h}))
" >> gpurun_out/r02_streamed_build_ncu.jsonl
done
tail -2 gpurun_out/plain_k4.log | cut -c1-300; cat gpurun_out/streamed_build.log; cat gpurun_out/r02_streamed_build_ncu.jsonl; head -12 gpurun_out/r02_align_fast_ncu_v2_lines.txt
