import json,sys
d=json.load(open(sys.argv[1]))
print('value %.2f it/s  ms/step %.3f (wall %.3f)  e2e %.2f it/s (%.2f ms)  stock e2e %.2f ms  cpu %s  ref_cuda %s' % (
    d['value'], d['ms_per_step'], d.get('wall_ms_per_step', 0), d['e2e']['value'], d['e2e']['ms_per_step'],
    (d.get('e2e_stock_api') or {}).get('ms_per_step', 0), d['cpu_baseline'] and round(d['cpu_baseline']['value'],4), d.get('reference_cuda')))
print({k:round(v['ms_per_step'],3) for k,v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])})
r=d['roofline']; print(r['kernel'], r['bound'], 'frac', round(r['frac'],4), 'two-pipe kernel %.3f step %.3f' % (r.get('kernel_two_pipe_frac',0), r.get('step_two_pipe_frac',0)), 'clocks', d['clocks'])
if d.get('strong'): print(json.dumps(d['strong'], indent=1))
