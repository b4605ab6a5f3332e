"""tcgen05 linear layers with fused epilogues vs a plain torch fp32 reference of the same op on
the same bf16-rounded inputs (bf16 output: 2e-2 of the output scale)."""
import pytest
import torch

from helpers import pkg

pytestmark = pytest.mark.gpu


def _close(a, b, tol=2e-2):
    a, b = a.float(), b.float()
    scale = b.abs().max().item() + 1e-6
    err = (a - b).abs().max().item()
    assert err <= tol * scale, (err, scale)


def _rand(shape, gen, std=1.0):
    return (torch.randn(*shape, generator=gen) * std).cuda().to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,relu,with_bias", [(128, 128, 64, False, True), (256, 512, 512, True, True), (1000, 2048, 512, True, True),
                                                  (77, 512, 2048, False, False), (10688, 2048, 512, True, True), (1, 128, 128, True, True)])
def test_linear_act(M, N, K, relu, with_bias):
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(M + N + K)
    x, w = _rand((M, K), gen), _rand((N, K), gen, K ** -0.5)
    b = torch.randn(N, generator=gen).cuda() if with_bias else None
    y = ops.linear_act(x, w, b, relu=relu)
    ref = torch.nn.functional.linear(x.float(), w.float(), b)
    if relu:
        ref = torch.relu(ref)
    assert y.dtype == torch.bfloat16 and y.shape == (M, N)
    _close(y, ref)


def test_linear_act_leading_dims_and_errors():
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(3)
    x, w = _rand((3, 50, 512), gen), _rand((512, 512), gen, 0.05)
    y = ops.linear_act(x, w, None)
    assert y.shape == (3, 50, 512)
    _close(y, torch.nn.functional.linear(x.float(), w.float()))
    with pytest.raises(RuntimeError):
        ops.linear_act(_rand((4, 64), gen), _rand((100, 64), gen))          # N not a multiple of 128
    with pytest.raises(RuntimeError):
        ops.linear_act(_rand((4, 80), gen), _rand((128, 80), gen))          # K not a multiple of 64
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        ops.linear_act(torch.zeros(4, 64, dtype=torch.bfloat16), w)          # CPU tensor: no fallback


@pytest.mark.parametrize("one_kernel", [False, True])
@pytest.mark.parametrize("M,K", [(128, 512), (300, 2048), (10688, 2048), (5, 64), (129, 512)])
def test_linear_residual_layernorm(M, K, one_kernel):
    """Both routes: the persistent GEMM + the bf16 LayerNorm kernel (default), and the single kernel that keeps the fp32 row
    in tensor memory."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(M + K)
    x, w, res = _rand((M, K), gen), _rand((512, K), gen, K ** -0.5), _rand((M, 512), gen)
    b = torch.randn(512, generator=gen).cuda()
    g = (1.0 + 0.1 * torch.randn(512, generator=gen)).cuda()
    be = (0.1 * torch.randn(512, generator=gen)).cuda()
    y = ops.linear_residual_layernorm(x, w, b, res, g, be, eps=1e-5, one_kernel=one_kernel)
    ref = torch.nn.functional.layer_norm(torch.nn.functional.linear(x.float(), w.float(), b) + res.float(), (512,), g, be, 1e-5)
    assert y.dtype == torch.bfloat16 and y.shape == (M, 512)
    _close(y, ref)


def test_feed_forward_block_matches_reference_module_math():
    """PositionwiseFeedForward.forward (module.py:46-53) with dropout off: w_2(relu(w_1(x))) + residual -> LayerNorm."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(9)
    B, T, d, di = 4, 167, 512, 2048
    x = _rand((B, T, d), gen)
    w1, b1 = _rand((di, d), gen, d ** -0.5), torch.randn(di, generator=gen).cuda() * 0.1
    w2, b2 = _rand((d, di), gen, di ** -0.5), torch.randn(d, generator=gen).cuda() * 0.1
    g, be = torch.ones(d).cuda(), torch.zeros(d).cuda()
    h = ops.linear_act(x, w1, b1, relu=True)
    y = ops.linear_residual_layernorm(h, w2, b2, x, g, be)
    hr = torch.relu(torch.nn.functional.linear(x.float(), w1.float(), b1)).to(torch.bfloat16).float()   # the bf16 hand-over
    ref = torch.nn.functional.layer_norm(torch.nn.functional.linear(hr, w2.float(), b2) + x.float(), (d,), g, be, 1e-5)
    _close(y, ref)
    with pytest.raises(RuntimeError):      # the single kernel is built for d_model = 512
        ops.linear_residual_layernorm(h, _rand((256, di), gen), None, _rand((B, T, 256), gen), torch.ones(256).cuda(), torch.zeros(256).cuda(),
                                      one_kernel=True)
    w3, r3 = _rand((256, di), gen, di ** -0.5), _rand((B, T, 256), gen)
    y3 = ops.linear_residual_layernorm(h, w3, None, r3, torch.ones(256).cuda(), torch.zeros(256).cuda())      # the two-kernel route: 256 / 512 / 1024
    _close(y3, torch.nn.functional.layer_norm(torch.nn.functional.linear(h.float(), w3.float()) + r3.float(), (256,), None, None, 1e-5))


def test_modules_take_the_fused_path_in_bf16_eval():
    """PositionwiseFeedForward / MultiheadAttention in eval + no_grad + bf16 run the fused kernels and agree with their
    own unfused forward (same bf16 parameters, torch ops) to bf16 accuracy."""
    module = pkg("transformer.module")
    attention = pkg("transformer.attention")
    torch.manual_seed(5)
    ffn = module.PositionwiseFeedForward(512, 2048, dropout=0.1).cuda().bfloat16().eval()
    mha = attention.MultiheadAttention(512, 8, dropout=0.1).cuda().bfloat16().eval()
    x = torch.randn(3, 167, 512, device="cuda").bfloat16()
    lib = pkg("_lib")
    with torch.no_grad():
        n0 = lib.launch_count()
        y_f = ffn(x)
        assert lib.launch_count() - n0 == 3                     # relu(w_1 x + b); w_2 h + b; LayerNorm(. + x)
        o_f, _ = mha(x, x, x)
    y_u = ffn(x)                                                 # grad mode: the torch path
    o_u, _ = mha(x, x, x)
    _close(y_f, y_u.detach(), 3e-2)
    _close(o_f, o_u.detach(), 3e-2)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 512, 512), (1344, 4233, 512), (77, 2048, 512), (1000, 512, 2048), (5, 36, 20)])
def test_linear_f32_three_tf32_products(M, N, K):
    """fp32 in, fp32 out, tensor cores: against an fp64 reference the error must be at fp32 level (a single TF32
    product would be ~1e-3), and no worse than a few times torch's own fp32 (SIMT) GEMM."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, generator=gen).cuda()
    w = (torch.randn(N, K, generator=gen) * K ** -0.5).cuda()
    b = torch.randn(N, generator=gen).cuda()
    y = ops.linear_f32(x, w, b)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    scale = ref.abs().max().item()
    err = (y.double() - ref).abs().max().item()
    err_torch = (torch.nn.functional.linear(x, w, b).double() - ref).abs().max().item()
    assert y.dtype == torch.float32 and y.shape == (M, N)
    # measured on B200: 4e-6 of the output scale at K = 512, 1.3e-5 at K = 2048 (the tensor core's fp32 accumulation
    # truncates, so the error grows with K; torch's SIMT fp32 GEMM: 7e-7 / 1.6e-6)
    assert err <= 1e-5 * scale * max(1.0, K / 512.0), (err, scale, err_torch)
    assert err <= 16 * err_torch + 1e-6 * scale, (err, err_torch)


@pytest.mark.parametrize("shape,N,K,bias", [((4, 21, 512), 2048, 512, True), ((1344, 512), 4233, 512, False), ((3, 7, 320), 512, 320, True),
                                            ((5, 5, 64), 36, 64, True)])
def test_linear_f32_autograd(shape, N, K, bias):
    """Forward, dx, dW, db of the differentiable fp32 tensor-core linear layer against torch in fp64."""
    ops = pkg("ops")
    gen = torch.Generator().manual_seed(N + K)
    x = torch.randn(*shape, generator=gen).cuda().requires_grad_(True)
    w = (torch.randn(N, K, generator=gen) * K ** -0.5).cuda().requires_grad_(True)
    b = torch.randn(N, generator=gen).cuda().requires_grad_(True) if bias else None
    gy = torch.randn(*shape[:-1], N, generator=gen).cuda()
    y = ops.linear_f32_autograd(x, w, b)
    y.backward(gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(gy.double())
    for got, ref in ((y, yd), (x.grad, xd.grad), (w.grad, wd.grad)) + (((b.grad, bd.grad),) if bias else ()):
        scale = ref.abs().max().item()
        assert (got.double() - ref.detach()).abs().max().item() <= 2e-5 * scale, (got.shape,)


def test_shell_linear_runs_on_the_tensor_core_kernel_and_can_be_switched_off(monkeypatch):
    module = pkg("transformer.module")
    lib = pkg("_lib")
    torch.manual_seed(2)
    lin = module.Linear(512, 2048).cuda()
    x = torch.randn(6, 21, 512, device="cuda")
    monkeypatch.setattr(module, "USE_TENSOR_CORE_FP32", False)
    n0 = lib.launch_count()
    y_ref = lin(x)                                   # torch's F.linear (cuBLAS SIMT sgemm)
    assert lib.launch_count() == n0
    monkeypatch.setattr(module, "USE_TENSOR_CORE_FP32", True)
    n0 = lib.launch_count()
    y = lin(x)                                       # the default: csrc/gemm2.cu
    assert 1 <= lib.launch_count() - n0 <= 2          # the GEMM (+ its split-K reduction)
    assert (y - y_ref).abs().max().item() <= 2e-5 * y_ref.abs().max().item()
    assert sorted(lin.state_dict().keys()) == ["bias", "weight"]
