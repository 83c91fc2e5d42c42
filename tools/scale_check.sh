# what the driver does for SCALE: the bench at N = 1 and N = 8 with --steps 20 --warmup 5 (one 8-GPU box)
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_n8_$i.json 2> gpurun_out/r02_scale_n8_$i.err
done
python - <<'PY'
import json
def load(p):
    for l in open(p):
        if l.startswith("{"): return json.loads(l)
a = load("gpurun_out/r02_scale_n1.json")
print("N=1", a["value"], a["ms_per_step"], a["roofline"]["kernel_ms"], a["e2e"]["value"], a["e2e"]["factored"]["value"])
for i in (1, 2):
    b = load(f"gpurun_out/r02_scale_n8_{i}.json")
    print("N=8", b["value"], b["ms_per_step"], b["roofline"]["kernel_ms"], "eff", b["value"] / (8 * a["value"]), b["arm"]["parallelism"][-90:], b["e2e"]["value"], b["e2e"]["factored"]["value"], "fact eff", b["e2e"]["factored"]["value"] / (8 * a["e2e"]["factored"]["value"]))
    print({k: (v["value"], v["value"] / (8 * a["workloads"][k]["value"])) for k, v in b["workloads"].items()})
PY
