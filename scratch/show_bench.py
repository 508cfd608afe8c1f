import json, sys
d=json.load(open(sys.argv[1]))
print("ms/step", round(d['ms_per_step'],3), "utt/s %.1fM"%(d['value']/1e6), "e2e", d['e2e'].get('ms_per_step'), "share", d.get('kernel_time_share_of_step'), d.get('clocks'))
tot=0
for k,v in d['kernels'].items():
    t=v['calls_per_step']*v['avg_ms']; tot+=t
    print("%-32s x%-3g %7.4f ms  tot %7.4f  frac %s"%(k, v['calls_per_step'], v['avg_ms'], t, v.get('frac_of_hbm_peak')))
print("sum", tot)
