import json, sys
l = open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.log').read().strip().splitlines()[-1]
d = json.loads(l)
ks = d.pop('kernel_shares', None) or {}
print('value %.1f %s  ms/step %.2f | e2e %.1f (%.2f ms) | roofline %s | cpu %s | launches %s | clocks %s' % (
    d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'],
    {k: d['roofline'][k] for k in ('achieved', 'frac', 'share_of_step')} if d.get('roofline') else None,
    d['cpu_baseline'] and round(d['cpu_baseline']['value'], 2), d['gpu_launches'], d['clocks']))
for k, v in ks.items():
    print('%-32s %3d %9.3f ms %6.3f' % (k, v['launches_per_step'], v['ms_per_step'], v['share']))
