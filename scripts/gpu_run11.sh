mkdir -p gpurun_out
rm -f gpurun_out/summary11.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "attention or pv_probe" > gpurun_out/tests11.log 2>&1; echo "attn tests exit $?" >> gpurun_out/summary11.txt
timeout 120 python - > gpurun_out/attn_bench11.log 2>&1 <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from deer_vla_b200.engine import op_attention
B,H,S=448,16,257
q=torch.randn(B,S,H,64,device='cuda').bfloat16(); k=torch.randn(B,S,H,64,device='cuda').bfloat16(); v=torch.randn(B,S,H,64,device='cuda').bfloat16()
for impl in ('tc','flash'):
    for _ in range(2): op_attention(q,k,v,0.125,impl=impl)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): op_attention(q,k,v,0.125,impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(impl, 'ms', round(ms,3), 'TFLOP/s', round(4*B*H*S*S*64/ms/1e9,1))
PY
echo "attn bench exit $?" >> gpurun_out/summary11.txt
cat gpurun_out/summary11.txt; tail -15 gpurun_out/tests11.log; cat gpurun_out/attn_bench11.log
