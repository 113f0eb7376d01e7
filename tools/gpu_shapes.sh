# other BASELINE shapes on the resident leg: 1KGP3 (5,008 haplotypes, 220 blocks) and biobank (1M haplotypes, 2 blocks)
mkdir -p gpurun_out
T=${T:-r01h}
timeout 600 python bench.py --samples 2504 --blocks 220 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_kgp.json 2> gpurun_out/${T}_bench_kgp.err; echo "kgp rc=$?"
timeout 900 python bench.py --samples 500000 --blocks 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_biobank.json 2> gpurun_out/${T}_bench_biobank.err; echo "biobank rc=$?"
tail -3 gpurun_out/${T}_bench_kgp.err gpurun_out/${T}_bench_biobank.err
python - <<P
import json
for n in ("kgp","biobank"):
    try:
        d=json.loads(open('gpurun_out/${T}_bench_%s.json'%n).read().strip().splitlines()[-1]); k=d["kernels"]
        print(n, "value %.1f enc %.1f dec %.1f verified %s | " % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.3), d["call_wall_ms_per_step"], d["config"])
    except Exception as e:
        print(n, "failed", e)
P
