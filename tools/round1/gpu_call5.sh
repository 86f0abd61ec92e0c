#!/bin/bash
mkdir -p gpurun_out
run() { # name, extra args
  local out=gpurun_out/r5_$1.json
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $2 > $out 2>gpurun_out/r5_$1.err
  python - "$1" $out <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], round(d['value']/1e9,3),'Gsteps/s frac',round(d['roofline']['frac'],3), d['roofline']['kernel'], 'ms', round(d['ms_per_step'],2), flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e, open(sys.argv[2].replace('.json','.err')).read()[-300:])
PY
}
python -m pytest tests/test_gpu_ingest.py tests/test_gpu_parity.py -q -x > gpurun_out/t5.log 2>&1; echo "tests: $(tail -1 gpurun_out/t5.log)"
run pl ""
run er "--workload er-100k-1M-sparseotf"
run er_g16 "--workload er-100k-1M-sparseotf --flags $((16<<8))"
run plw "--workload powerlaw-1M-10M-sparseotf-weighted"
python - <<'PY' 2>&1 | tail -8
import time, numpy as np, torch, os
from pecanpy_b200.engine import WalkEngine
from pecanpy_b200 import synth
from pecanpy_b200.ingest import csr_from_edges_device
z=np.load('/tmp/b2w_bench_cache/powerlaw_1000000_10000000_1_1.npz')
eng=WalkEngine.from_csr(z['indptr'],z['indices'],z['data'])
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); thr=eng.compute_thresholds(0.5); torch.cuda.synchronize(); dt=time.perf_counter()-t
print('thresholds csr 1M nodes / 2e7 weights: %.2f ms'%(dt*1e3))
t=time.perf_counter()
indptr=z['indptr'].astype(np.int64); data=z['data']
for i in range(20000):
    row=data[indptr[i]:indptr[i+1]]
    if row.size: row.mean()+0.5*row.std()
print('reference-style numpy loop: %.1f us/node'%((time.perf_counter()-t)/20000*1e6))
eng.close()
# device CSR build at config #3 size: edges back out of the CSR (upper triangle), shuffled
ip=z['indptr'].astype(np.int64); rows=np.repeat(np.arange(ip.size-1,dtype=np.uint32), np.diff(ip)); cols=z['indices']
keep=rows<cols; src=rows[keep]; dst=cols[keep]; w=z['data'][keep].astype(np.float64)
perm=np.random.default_rng(0).permutation(src.size); src,dst,w=src[perm],dst[perm],w[perm]
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    a,b,c,k=csr_from_edges_device(ip.size-1,src,dst,w,False,return_tensors=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print('device CSR build, %d undirected edges: %.1f ms (incl. H2D of the edge arrays); nnz=%d'%(src.size,dt*1e3,k))
ok=np.array_equal(a.cpu().numpy().view(np.uint32),z['indptr']) and np.array_equal(b[:k].cpu().numpy().view(np.uint32),z['indices']) and np.array_equal(c.cpu().numpy(),z['data'])
print('device CSR equals the generator CSR:',ok)
t=time.perf_counter()
from pecanpy_b200.graph import _coo_to_csr
r2=np.concatenate([src,dst]).astype(np.int64); c2=np.concatenate([dst,src]).astype(np.int64); w2=np.concatenate([w,w])
_coo_to_csr(ip.size-1,r2,c2,w2); print('host lexsort CSR build: %.1f s'%(time.perf_counter()-t))
rng=np.random.default_rng(1); n=8192
nz=rng.random((n,n))<0.3; dd=np.where(nz,0.01+0.99*rng.random((n,n)),0.0)
e2=WalkEngine.from_dense(dd,nz)
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); e2.compute_thresholds(0.5); torch.cuda.synchronize(); dt=time.perf_counter()-t
print('thresholds dense n=8192 (30%% density): %.2f ms'%(dt*1e3))
PY
