import json, sys
r=json.load(open('/root/repo/gpurun_out/probe.json'))
key = sys.argv[1] if len(sys.argv)>1 else 'net_B4'
L=r[key]['layers']
print('sum layers ms', r[key]['sum_layers_ms'], 'step', r[key]['ms_per_step'])
cats={}
for n,t,f in L:
    if 'chain' in n: c='fused chain'
    elif 'proj' in n and 'block' in n: c='conv3x3(RB)'
    elif n.endswith('.norm') : c='gn_apply'
    elif 'norm2' in n: c='layernorm'
    elif 'ff.net' in n or 'proj_out' in n: c='attn gemm'
    elif 'res_conv' in n: c='res_conv'
    elif n=='init_conv': c='init_conv'
    elif 'nearest' in n: c='upsample'
    elif n.endswith('.3.1') or n.endswith('.3'): c='down/up conv'
    elif 'fc' in n: c='shot mlp'
    else: c='other'
    cats[c]=cats.get(c,0)+t
for c,t in sorted(cats.items(), key=lambda x:-x[1]): print(f"{c:16s} {t:8.1f} us")
thr = float(sys.argv[2]) if len(sys.argv)>2 else 30
for n,t,f in L:
    if t>thr: print(f"{n:34s} {t:8.1f} us  {f:7.1f} TF")
