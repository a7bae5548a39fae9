import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    if 'cfg' in d:
        print(d['cfg'], end='  ')
    else:
        r = d['roofline']
        print(f"rows {d['config']['rows']} qps {d['value']:.0f} ms {d['ms_per_step']:.3f} scan_us {r['kernel_us']:.0f} {r['bound']} frac {r['frac']:.3f} hbm {r['hbm_gbs_of_scan']:.0f} clk {d['clocks']['sm_mhz']} e2e {d['e2e']['value']:.0f}")
