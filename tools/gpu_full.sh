python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest.txt; cat gpurun_out/r2_pytest.txt; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 800 gpurun_out/r2_bench.err; python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; python - <<EOP
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d.get("parity_check"), d.get("cpu_baseline",{}).get("value"))
print(json.dumps(d.get("pmp_build"))[:1500])
r=json.load(open("gpurun_out/r2_bench_reference.json")); print(r["value"], json.dumps(r.get("pmp_build"))[:600])
EOP
