"""Pretty-prints a bench.py JSON line (developer tool)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.json').read().splitlines() if l.startswith('{')][-1])
print(f"value {d['value']:.2f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.2f}  launches {d['gpu_launches']}  clocks {d['clocks']}")
print('fp64 stage', {k: v for k, v in d['fp64_tensor_stage'].items() if k != 'note'})
r = d['roofline']; print('roofline', r['kernel'], r['achieved'], r['frac'])
for k, v in d['stages'].items():
    print(f"{k:20s} {v['ms_per_step']:9.3f} ms  {100*v['share']:5.1f}%  {v['tflops']:6.2f} TF/s {v['gbs']:7.0f} GB/s n={v['launches_per_step']:.0f}")
if d.get('cpu_baseline'): print(d['cpu_baseline'])
