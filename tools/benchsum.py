import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['kernels']
print('value {:.0f} ms {:.3f} e2e {:.0f} | fwd {} dgrad {} wgrad {} ewf {} ewb {}'.format(d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else 0, k['conv_fwd']['by_stage_ms'], k['conv_dgrad']['by_stage_ms'], k['conv_wgrad']['ms_per_step'], k['bn_relu_pool_fwd']['ms_per_step'], k['bn_relu_pool_bwd']['ms_per_step']))
