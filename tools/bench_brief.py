import json,sys
d=json.load(open(sys.argv[1]))
print('value %.2f it/s  ms/step %.3f  e2e %.2f it/s (%.2f ms)  cpu %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline'] and round(d['cpu_baseline']['value'],4)))
print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
r=d['roofline']; print(r['kernel'], r['bound'], 'frac', round(r['frac'],4), 'clocks', d['clocks'])
