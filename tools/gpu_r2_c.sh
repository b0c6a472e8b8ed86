#!/bin/bash
# round 2, third GPU pass: where does a batched step's time go?  launch lists (ncu, per kernel) for both weight sources,
# a --set full capture of the quantised-weight GEMM, step time vs batch size, then config 4 / 5 first numbers.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for B in 8 16 32 64; do timeout 200 python tools/batch_probe.py $B 576 8 2; done > gpurun_out/r2c_probe_q.log 2>&1
for B in 32 128; do timeout 200 python tools/batch_probe.py $B 576 8 1; done > gpurun_out/r2c_probe_w16.log 2>&1
cat gpurun_out/r2c_probe_q.log gpurun_out/r2c_probe_w16.log | grep "^{"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_q.csv python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2c_ncu_q.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_w16.csv python tools/batch_probe.py 32 576 1 1 > gpurun_out/r2c_ncu_w16.log 2>&1
python - <<'PY'
import csv, collections, sys
for f in ("gpurun_out/r2c_launches_q.csv", "gpurun_out/r2c_launches_w16.csv"):
    try:
        rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
        hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
        # keep the LAST 261 launches = one timed step
        body = rows[1:]
        agg = collections.OrderedDict()
        for r in body[-261:]:
            k = r[ki].split("(")[0][:60]; v = float(r[vi].replace(",", ""))
            a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
        tot = sum(a[1] for a in agg.values())
        print(f, "last step: %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]): print("  %-62s n=%4d  %9.1f us  (%.1f us each)" % (k, a[0], a[1] / 1e3, a[1] / 1e3 / a[0]))
    except Exception as ex:
        print(f, "failed:", ex)
PY
# one --set full capture of the quantised-weight GEMM (the gate/up launch of a middle layer) and of the batched attention
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qgemm_kernel -s 150 -c 1 -o gpurun_out/r2c_qgemm_full python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2c_ncu_full_q.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:batch_attn_kernel -s 40 -c 1 -o gpurun_out/r2c_battn_full python tools/batch_probe.py 32 576 1 2 > gpurun_out/r2c_ncu_full_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
( time GL_BENCH_WATCHDOG_S=120 timeout 900 python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu ) > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
tail -5 gpurun_out/r2c_bench_c4.err; cut -c1-1800 gpurun_out/r2c_bench_c4.json
( time GL_TC5_BN=128 GL_BENCH_WATCHDOG_S=120 timeout 600 python bench.py --workload config4 --steps 1 --warmup 1 --no-cpu ) > gpurun_out/r2c_bench_c4_bn128.json 2> gpurun_out/r2c_bench_c4_bn128.err
tail -3 gpurun_out/r2c_bench_c4_bn128.err; cut -c1-400 gpurun_out/r2c_bench_c4_bn128.json
( time GL_BENCH_WATCHDOG_S=120 timeout 600 python bench.py --workload config5 --steps 1 --warmup 1 --docs 640 --no-cpu ) > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err
tail -5 gpurun_out/r2c_bench_c5.err; cut -c1-1800 gpurun_out/r2c_bench_c5.json
