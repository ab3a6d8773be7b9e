import json,sys
d=json.load(open(sys.argv[1]))
print('value %.3g  ms/step %.3f  e2e %.4g  pageable %.4g'%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable_inputs"]["value"]))
r=d["roofline"]; print(r["kernel"], 'launch_us %.1f frac %.3f share %.2f'%(r["launch_us"], r["frac"], r['share_of_step']), 'in-region us', r['timed_region_overlapped']['launch_us'])
print(' alone us', r.get("kernel_us_alone"))
print(' path', d['roofline_path']['frac'], 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'])
for k in ("kitti360_long_horizon","highres_1024"):
    if k in d.get("extra",{}): print(k, round(d["extra"][k]["rasterise_ms_per_bev"],4), d["extra"][k]["kernel_us_per_bev"])
