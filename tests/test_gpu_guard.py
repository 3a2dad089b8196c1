"""Guard-band ("canary") bounds checks of every kernel family through the C ABI.

compute-sanitizer is not available on the GPU pool, so this is the substitute: each case runs an op twice --
once on ordinary tensors, once with EVERY input placed between two 4 KiB bands of NaN bit patterns (0xFF) and EVERY
output (and workspace) allocated between two bands of a sentinel (0xA5) -- at ragged shapes that exercise the M / N / K
tails, partial query / key tiles and the last chunks of the persistent loops.  It then asserts
  * the sentinel bands around every output are untouched (no out-of-bounds WRITE: TMA stores clipped by the tensor map,
    tail rows of the register-store paths, the per-warp chunk loops), and
  * the guarded run reproduces the plain run bit for bit (an out-of-bounds READ that reaches a result would carry the
    NaN bands into it; TMA loads beyond the tensor extent are zero-filled by the tensor map instead).
The outputs the wrappers allocate themselves are redirected into the arena by swapping the module-level ``torch`` name
of the shim modules for a proxy whose ``empty`` / ``empty_like`` / ``zeros`` come out of the arena."""
import contextlib
import math

import pytest
import torch

import pcd_b200 as P
from gpu_util import DEV, build_model, to_dev
from oracle import cases, det

pytestmark = pytest.mark.gpu
ops = P.ops
GUARD = 4096


class Arena:
    def __init__(self):
        self.blocks = []   # (base uint8 tensor, payload bytes, band byte)

    def _carve(self, nbytes, band):
        base = torch.full((GUARD + nbytes + GUARD,), band, dtype=torch.uint8, device=DEV)
        self.blocks.append((base, nbytes, band))
        return base[GUARD:GUARD + nbytes]

    def output(self, shape, dtype):
        """Uninitialised output / workspace between two 0xA5 bands (the payload starts as 0xA5 too)."""
        shape = tuple(shape)
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        return self._carve(n, 0xA5).view(dtype).view(shape)

    def place(self, t):
        """Copy of an input between two NaN (0xFF) bands."""
        if t is None:
            return None
        t = t.contiguous()
        v = self._carve(t.numel() * t.element_size(), 0xFF).view(t.dtype).view(t.shape)
        v.copy_(t)
        return v

    def check(self):
        torch.cuda.synchronize()
        for i, (base, n, band) in enumerate(self.blocks):
            lo, hi = base[:GUARD], base[GUARD + n:]
            assert bool((lo == band).all()), f"block {i} ({n} B): bytes BEFORE the buffer were overwritten"
            assert bool((hi == band).all()), (f"block {i} ({n} B): bytes AFTER the buffer were overwritten "
                                              f"(first at +{int((hi != band).nonzero()[0])})")


class GuardedTorch:
    """Stands in for the ``torch`` module inside the shim modules: allocations come out of the arena."""

    def __init__(self, arena):
        self._arena = arena

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _shape(args):
        return tuple(args[0]) if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)) else tuple(args)

    def empty(self, *args, dtype=None, device=None, **kw):
        return self._arena.output(self._shape(args), dtype or torch.float32)

    def empty_like(self, t, **kw):
        return self._arena.output(t.shape, kw.get("dtype", t.dtype))

    def zeros(self, *args, dtype=None, device=None, **kw):
        return self._arena.output(self._shape(args), dtype or torch.float32).zero_()


@contextlib.contextmanager
def guarded_allocations(arena, *modules):
    saved = []
    proxy = GuardedTorch(arena)
    for m in modules:
        for name in ("torch", "th"):
            if getattr(m, name, None) is torch:
                saved.append((m, name))
                setattr(m, name, proxy)
    try:
        yield
    finally:
        for m, name in saved:
            setattr(m, name, torch)


def run_both(fn, inputs, modules=(ops,)):
    """fn(*inputs) on plain tensors, then on guarded copies with guarded outputs; returns (plain, guarded, arena)."""
    plain = fn(*inputs)
    torch.cuda.synchronize()
    arena = Arena()
    guarded_in = [arena.place(t) if torch.is_tensor(t) else t for t in inputs]
    with guarded_allocations(arena, *modules):
        got = fn(*guarded_in)
    arena.check()
    return plain, got, arena


def same(a, b):
    if torch.is_tensor(a):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert torch.equal(a, b) or torch.equal(a.view(torch.uint8), b.view(torch.uint8)), \
            f"guarded run differs from the plain run: max |d| = {float((a.float() - b.float()).abs().max())}"
    else:
        for x, y in zip(a, b):
            same(x, y)


def bf(t):
    return t.to(DEV).bfloat16()


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(389, 264, 200), (128 * 5 + 1, 512, 512), (77, 1536, 520), (1, 8, 8)])
@pytest.mark.parametrize("epi", ["bias", "gelu", "residual"])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_gemm_bf16_tails(M, N, K, epi, out_dtype):
    if epi == "residual" and out_dtype != torch.float32:
        pytest.skip("the residual epilogue writes fp32 (pcd_gemm_bf16 contract)")
    a, w = bf(det.normal((M, K), 9001)), bf(det.uniform((N, K), 9002, 1 / math.sqrt(K)))
    b = det.uniform((N,), 9003, 0.5).to(DEV)
    r = det.normal((M, N), 9004).to(DEV) if epi == "residual" else None

    def fn(a, w, b, r):
        return ops.linear(a, w, b, epilogue=P._lib.EPI_BIAS_GELU if epi == "gelu" else P._lib.EPI_BIAS, residual=r,
                          out_dtype=out_dtype)
    plain, got, _ = run_both(fn, (a, w, b, r))
    same(plain, got)


@pytest.mark.parametrize("M,N,K", [(77, 50, 36), (300, 129, 68), (1, 1, 4)])   # K % 4 == 0 is the kernel's contract
def test_gemm_f32_tails(M, N, K):
    a, w = det.normal((M, K), 9011).to(DEV), det.uniform((N, K), 9012, 0.2).to(DEV)
    b = det.uniform((N,), 9013, 0.5).to(DEV)
    plain, got, _ = run_both(lambda a, w, b: ops.linear(a, w, b, epilogue=P._lib.EPI_BIAS_GELU), (a, w, b))
    same(plain, got)


@pytest.mark.parametrize("M,N,K", [(517, 512, 200), (128 * 9 + 3, 256, 1024)])   # contract: N % 256 == 0, M >= 512
def test_gemm_residual_stats_tails(M, N, K):
    """In-place fp32 stream update + bf16 copy + row statistics: three outputs, all guarded."""
    a, w = bf(det.normal((M, K), 9021)), bf(det.uniform((N, K), 9022, 1 / math.sqrt(K)))
    b = det.uniform((N,), 9023, 0.5).to(DEV)
    h0 = det.normal((M, N), 9024).to(DEV)

    def fn(a, w, b, h):
        h = h.clone() if h is h0 else h      # guarded copy is updated in place inside its bands
        hb, st = ops.linear_residual_stats(a, w, b, h)
        return h, hb, st
    plain, got, _ = run_both(fn, (a, w, b, h0))
    same(plain, got)


@pytest.mark.parametrize("M,N,K,gelu", [(513, 256, 512, False), (128 * 4 + 17, 2048, 512, True), (700, 512, 256, True)])
def test_gemm_layernorm_folded_tails(M, N, K, gelu):
    h = det.normal((M, K), 9031).to(DEV)
    w = bf(det.uniform((N, K), 9032, 1 / math.sqrt(K)))
    colsum, const = w.float().sum(dim=1).contiguous(), det.uniform((N,), 9033, 0.5).to(DEV)
    hb, st = ops.cast_rowstats(h)
    plain, got, _ = run_both(lambda hb, st, w, cs, c: ops.linear_layernorm_folded(hb, st, w, cs, c, gelu=gelu),
                             (hb, st, w, colsum, const))
    same(plain, got)


@pytest.mark.parametrize("variant", [0, 8, 11, 9, 5])
@pytest.mark.parametrize("B,H,L", [(3, 2, 1026), (2, 8, 130), (5, 2, 70), (1, 1, 1)])
def test_self_attention_bf16_tiles(variant, B, H, L):
    qkv = bf(det.normal((B, L, H * 192), 9041, std=1.5))
    plain, got, _ = run_both(lambda x: ops.self_attention(x, H, variant=variant), (qkv,))
    same(plain, got)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Lq,Lkv", [(130, 77), (1026, 5), (64, 257)])
def test_cross_attention_tiles(dtype, Lq, Lkv):
    B, H = 3, 2
    q = det.normal((B, Lq, H * 64), 9051).to(DEV).to(dtype)
    kv = det.normal((B, Lkv, H * 128), 9052).to(DEV).to(dtype)
    plain, got, _ = run_both(lambda q, kv: ops.cross_attention(q, kv, H), (q, kv))
    same(plain, got)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rotary_attention_tiles(dtype):
    B, H, L = 2, 2, 333
    qkv = det.normal((B, L, H * 192), 9061).to(DEV).to(dtype)
    coords = det.uniform((B, L, 3), 9062, 1.0).to(DEV)
    plain, got, _ = run_both(lambda x, c: ops.rotary_attention(x, c, H), (qkv, coords))
    same(plain, got)


@pytest.mark.parametrize("Lq,Lkv", [(131, 77), (643, 643), (1, 64)])
def test_attention_hd32_bf16_tiles(Lq, Lkv):
    """32-wide heads through 64-column TMA boxes: the store must drop the out-of-tensor half of every row."""
    B, H = 3, 4
    q, k, v = (bf(det.normal((B, n, H * 32), 9075 + i)) for i, n in enumerate((Lq, Lkv, Lkv)))
    plain, got, _ = run_both(lambda q, k, v: ops.attention_views(q, k, v, H, 0.42, 0.42, head_dim=32), (q, k, v))
    same(plain, got)


def test_attention_hd32_tiles():
    B, H, Lq, Lkv = 2, 4, 131, 77
    q, k, v = (det.normal((B, n, H * 32), 9070 + i).to(DEV) for i, n in enumerate((Lq, Lkv, Lkv)))
    plain, got, _ = run_both(lambda q, k, v: ops.attention_hd32(q, k, v, H), (q, k, v))
    same(plain, got)


@pytest.mark.parametrize("dim,c_in,n_prefix,n", [(512, 6, 2, 333), (512, 3, 0, 17), (384, 6, 1, 100), (1024, 6, 258, 40)])
def test_embed_and_output_proj_chunks(dim, c_in, n_prefix, n):
    """Register-resident and shared-memory kernels: the last (partial) 16-row chunk of the persistent warps."""
    seqs, x_seqs = 6, 3
    x = det.normal((x_seqs, c_in, n), 9081).to(DEV)
    w, b = det.uniform((dim, c_in), 9082, 0.4).to(DEV), det.uniform((dim,), 9083, 0.5).to(DEV)
    pre = det.normal((seqs, n_prefix, dim), 9084).to(DEV) if n_prefix else None
    g, beta = (1.0 + 0.1 * det.normal((dim,), 9085)).to(DEV), (0.1 * det.normal((dim,), 9086)).to(DEV)
    stats = dim % 128 == 0
    plain, got, _ = run_both(lambda x, w, b, pre, g, beta: ops.embed_tokens(x, w, b, pre, None, g, beta, seqs=seqs,
                                                                            with_stats=stats), (x, w, b, pre, g, beta))
    same(plain, got)
    h = plain[0] if stats else plain
    wo, bo = det.uniform((6, dim), 9087, 0.05).to(DEV), det.uniform((6,), 9088, 0.5).to(DEV)
    for y in (None, det.normal(tuple(h.shape), 9089).to(DEV)):
        p2, g2, _ = run_both(lambda h, g, beta, wo, bo, y: ops.output_proj(h, n_prefix, g, beta, wo, bo, y=y),
                             (h, g, beta, wo, bo, y))
        same(p2, g2)


@pytest.mark.parametrize("rows,dim", [(777, 512), (5, 128), (33, 1024)])
def test_rowwise_kernels(rows, dim):
    h = det.normal((rows, dim), 9091).to(DEV)
    y = bf(det.normal((rows, dim), 9092))
    g, beta = (1.0 + 0.1 * det.normal((dim,), 9093)).to(DEV), (0.1 * det.normal((dim,), 9094)).to(DEV)
    same(*run_both(lambda h: ops.cast_rowstats(h), (h,))[:2])
    for od in (torch.float32, torch.bfloat16):
        same(*run_both(lambda h, g, b: ops.layernorm(h, g, b, out_dtype=od), (h, g, beta))[:2])

    def add_ln(h_, y_, g_, b_):
        h_ = h_.clone() if h_ is h else h_
        return h_, ops.add_layernorm(h_, y_, g_, b_, out_dtype=torch.bfloat16)
    same(*run_both(add_ln, (h, y, g, beta))[:2])


@pytest.mark.parametrize("case,dtype", [("small_imagevec", torch.bfloat16), ("small_imagevec", torch.float32),
                                        ("small_upsample_grid", torch.bfloat16), ("full_imagevec", torch.bfloat16)])
def test_model_forward_workspace_and_output(case, dtype):
    """pcd_model_forward with its workspace carved to EXACTLY pcd_model_workspace_bytes between sentinel bands (every
    intermediate of every block lives in it) and a guarded output; same result as the ordinary forward."""
    model, cfg, _ = build_model(case, dtype)
    x, t, kw = cases.forward_inputs(case)
    x, t, kw = x.to(DEV), t.to(DEV), to_dev(kw)
    with torch.no_grad():
        plain = model(x, t, **kw).clone()
        torch.cuda.synchronize()
        model._ws.clear()
        model._prefix.clear()
        model._cond_key.clear()
        arena = Arena()
        with guarded_allocations(arena, P.transformer):
            got = model(arena.place(x), t, **kw)
        arena.check()
    assert len(arena.blocks) >= 3   # input, workspace, output (+ prefix tokens)
    same(plain, got)
    model._ws.clear()
    model._prefix.clear()
    model._cond_key.clear()


def test_sampler_and_ddpm_kernels():
    """Fused Heun predictor / corrector and the ancestral step on odd point counts (scalar tails of the vector loops)."""
    d = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
    B, C, N = 3, 6, 37
    x = det.normal((B, C, N), 9101).to(DEV)
    out = det.normal((B, 2 * C, N), 9102, std=0.7).to(DEV)
    noise = det.normal((B, C, N), 9103).to(DEV)
    t = torch.tensor([1023, 0, 500], device=DEV)
    model = lambda x_, t_, **kw: out

    def step(x, noise):
        r = d.p_sample(model, x, t, clip_denoised=True, noise=noise)
        return r["sample"], r["pred_xstart"]
    same(*run_both(step, (x, noise), modules=(P.gaussian_diffusion,))[:2])

    plan = P.HeunPlan(d, 8, 1e-3, 120.0, 7.0, 3.0)
    arena = Arena()
    with guarded_allocations(arena, P.k_diffusion):
        st = P.k_diffusion.HeunState(d, plan, (B, C, N), DEV, 3.0, True)
        mo = arena.place(det.normal((2 * B, 2 * C, N), 9104).to(DEV))
        nz = arena.place(noise)
        pred = arena.output((B, C, N), torch.float32)
        st.begin(nz)
        st.predictor(0, mo, pred)
        st.corrector(0, mo, nz)
    arena.check()
    assert torch.isfinite(st.x).all() and torch.isfinite(pred).all()


def test_point_cloud_kernels():
    a, b = det.normal((2, 301, 3), 9111).to(DEV), det.normal((2, 257, 3), 9112).to(DEV)
    same(*run_both(lambda a, b: ops.chamfer_distance_xyz(a, b), (a, b))[:2])
    same(*run_both(lambda a, b: ops.nearest_points(a, b), (a, b))[:2])
    pts = det.normal((2, 1000, 3), 9113).to(DEV)
    same(*run_both(lambda p: ops.farthest_point_sample(p, 65, 7), (pts,))[:2])
    big = det.normal((1, 9000, 3), 9114).to(DEV)     # > 8192 points: the global-workspace kernel
    same(*run_both(lambda p: ops.farthest_point_sample(p, 33, 0), (big,))[:2])
    same(*run_both(lambda a, b: ops.fscore_point_cloud_batch(a, b), (a, b))[:2])
