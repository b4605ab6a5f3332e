import sys, torch, importlib
sys.path.insert(0, ".")
ops = importlib.import_module("end-to-end_asr_pytorch_b200.ops"); lib = importlib.import_module("end-to-end_asr_pytorch_b200._lib")
F = torch.nn.functional
g = torch.Generator(device="cuda").manual_seed(1)
M, N, K = 102400, 2048, 512
x = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
ref = F.linear(x[:4096].double(), w.double())
for raw in (128, 256):
    lib.set_option("gemm_f32_bn", raw)
    for _ in range(2): ops.linear_f32(x, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = ops.linear_f32(x, w)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    err = ((y[:4096].double() - ref).abs().max() / ref.abs().max()).item()
    print("BN=%d: %.3f ms %.0f TFLOP/s err %.2e" % (raw, ms, 2.0 * M * N * K / ms / 1e9, err))
