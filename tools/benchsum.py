import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d.get('kernels') or {}
fam={n:(v['ms_per_step'], v.get('tflops')) for n,v in k.items() if isinstance(v,dict) and 'ms_per_step' in v}
print('value {:.0f} ms {:.3f} mhz {} e2e {:.0f} frac {} | {}'.format(d['value'], d['ms_per_step'], (d.get('clocks') or {}).get('sm_mhz'), d['e2e']['value'] if d.get('e2e') else 0, (d.get('roofline') or {}).get('frac'), fam))
