import importlib, sys, torch
sys.path.insert(0, "/root/repo")
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops"); lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
g = torch.Generator(device="cuda").manual_seed(1)
M, d, di = 102400, 512, 2048
x = torch.randn(M, d, device="cuda", generator=g).bfloat16()
w1 = (torch.randn(di, d, device="cuda", generator=g) * d ** -0.5).bfloat16(); b1 = torch.randn(di, device="cuda", generator=g)
w2 = (torch.randn(d, di, device="cuda", generator=g) * di ** -0.5).bfloat16(); b2 = torch.randn(d, device="cuda", generator=g)
h = ops.linear_act(x, w1, b1, relu=True)
flop = 2.0 * M * d * di
a = timed(lambda: ops.linear_act(x, w1, b1, relu=True)); print("linear_act w_1: %.3f ms %.0f TFLOP/s" % (a, flop / a / 1e9))
for var in (0, 1, 2):
    lib.set_option("gemm_variant", var)
    a = timed(lambda: ops.gemm_bf16(x, w1, bias=b1, relu=True)); print("gemm_bf16 w_1 variant %d: %.3f ms %.0f TFLOP/s" % (var, a, flop / a / 1e9))
    a = timed(lambda: ops.gemm_bf16(h, w2, bias=b2)); print("gemm_bf16 w_2 variant %d: %.3f ms %.0f TFLOP/s" % (var, a, flop / a / 1e9))
lib.set_option("gemm_variant", 0)
y1 = ops.linear_act(x, w1, b1, relu=True); y2 = ops.gemm_bf16(x, w1, bias=b1, relu=True)
print("equal:", torch.equal(y1, y2), (y1.float() - y2.float()).abs().max().item())
