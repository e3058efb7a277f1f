"""GPU bring-up probe for vl_gemm_bf16: correctness across operand majors / epilogues, a
descriptor sweep when an MN-major default fails, and timing of the ViT-L/14 shapes.
Run under gpurun:  python tools/probe_gemm.py > gpurun_out/probe_gemm.log 2>&1"""
import itertools
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def ref_gelu(x, quick):
    return x * torch.sigmoid(1.702 * x) if quick else torch.nn.functional.gelu(x)


def run(M, N, K, a_mn=False, b_mn=False, epi=L.EPI_LINEAR, bias=False, f32=False, split_k=1, quick=False, alpha=1.0, tag=""):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    Ag = A.t().contiguous() if a_mn else A
    Bg = B.t().contiguous() if b_mn else B
    bias_t = torch.randn(N, device=dev) if bias else None
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (L.EPI_RESIDUAL, L.EPI_GELU_BWD) else None
    aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == L.EPI_GELU else None
    D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    L.gemm(Ag, Bg, D, M=M, N=N, K=K, lda=Ag.shape[1], ldb=Bg.shape[1], ldd=N, a_mn=a_mn, b_mn=b_mn, epilogue=epi,
           bias=bias_t, aux_in=aux, aux_out=aux_out, ldaux=N, alpha=alpha, accumulate=(split_k > 1), split_k=split_k,
           act_quick=quick)
    torch.cuda.synchronize()
    acc = (A.float() @ B.float().t()) * alpha
    if epi == L.EPI_GELU_BWD:
        x = aux.float()
        xr = x.clone().requires_grad_(True)
        g = torch.autograd.grad(ref_gelu(xr, quick).sum(), xr)[0]
        ref = acc * g
    else:
        if bias:
            acc = acc + bias_t
        if epi == L.EPI_GELU:
            u = acc
            ref = ref_gelu(u, quick)
        elif epi == L.EPI_RESIDUAL:
            ref = acc + aux.float()
        else:
            ref = acc
    err = (D.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= 2e-2 * scale + 1e-3
    extra = ""
    if epi == L.EPI_GELU:
        e2 = (aux_out.float() - u).abs().max().item()
        extra = f" aux_err={e2:.3e}"
        ok = ok and e2 <= 2e-2 * u.abs().max().item()
    print(f"[{'OK ' if ok else 'BAD'}] {tag} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bias={int(bias)} "
          f"f32={int(f32)} split={split_k} quick={int(quick)} err={err:.3e} scale={scale:.2e}{extra}", flush=True)
    if not ok:
        bad = ((D.float() - ref).abs() > 2e-2 * scale + 1e-3)
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"      bad frac={bad.float().mean().item():.4f} rows[{rows.numel()}] {rows[:8].tolist()}.. cols[{cols.numel()}] {cols[:8].tolist()}..", flush=True)
    return ok


def sweep(which):
    print(f"--- descriptor sweep for {which}_mn", flush=True)
    base = 1 if which == "a" else 4
    for lbo, sbo, kstep in itertools.product([8192, 1024, 128, 16, 2048], [1024, 8192, 128, 2048], [2048, 32, 256, 4096]):
        L.debug_set(base, lbo)
        L.debug_set(base + 1, sbo)
        L.debug_set(base + 2, kstep)
        ok = run(128, 128, 64, a_mn=(which == "a"), b_mn=(which == "b"), f32=True, tag=f"sweep lbo={lbo} sbo={sbo} kstep={kstep}")
        if ok:
            print(f"*** {which}_mn works with lbo={lbo} sbo={sbo} kstep={kstep}", flush=True)
    for k in range(base, base + 3):
        L.debug_set(k, 0)


def timeit(M, N, K, a_mn=False, b_mn=False, epi=L.EPI_LINEAR, f32=False, split_k=1, iters=10, tag=""):
    A = torch.randn(K, M, device=dev).bfloat16() if a_mn else torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(K, N, device=dev).bfloat16() if b_mn else torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (L.EPI_RESIDUAL, L.EPI_GELU_BWD) else None
    aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == L.EPI_GELU else None
    D = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)

    def call():
        L.gemm(A, B, D, M=M, N=N, K=K, lda=A.shape[1], ldb=B.shape[1], ldd=N, a_mn=a_mn, b_mn=b_mn, epilogue=epi,
               bias=bias if epi != L.EPI_GELU_BWD else None, aux_in=aux, aux_out=aux_out, ldaux=N,
               accumulate=(split_k > 1), split_k=split_k)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # torch / cuBLAS comparison
    Aq = A.t() if a_mn else A
    Bq = B if b_mn else B.t()
    for _ in range(3):
        torch.matmul(Aq, Bq)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(Aq, Bq)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"[time] {tag} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} split={split_k}: {ms:.3f} ms  {tf:.0f} TFLOP/s"
          f"   (cuBLAS {ms2:.3f} ms {2.0 * M * N * K / ms2 / 1e9:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    if "--pair" in sys.argv:
        L.debug_set(8, 2)
        print("### CTA-pair (cta_group::2) kernel forced", flush=True)
    t0 = time.time()
    ok = run(128, 128, 64, f32=True, tag="smallest")
    ok &= run(256, 256, 128, tag="small")
    ok &= run(1000, 384, 1024, bias=True, tag="tail-M")
    ok &= run(300, 72, 200, bias=True, f32=True, tag="odd")
    okb = run(256, 256, 128, b_mn=True, f32=True, tag="b_mn")
    oka = run(256, 256, 128, a_mn=True, f32=True, tag="a_mn")
    if not okb:
        sweep("b")
    if not oka:
        sweep("a")
    run(384, 512, 1000, a_mn=True, b_mn=True, f32=True, tag="wgrad")
    run(384, 512, 4096, a_mn=True, b_mn=True, f32=True, split_k=4, tag="wgrad-splitk")
    run(520, 1024, 256, epi=L.EPI_GELU, bias=True, tag="gelu")
    run(520, 1024, 256, epi=L.EPI_GELU, bias=True, quick=True, tag="quickgelu")
    run(520, 1024, 256, epi=L.EPI_RESIDUAL, bias=True, tag="residual")
    run(520, 1024, 256, epi=L.EPI_GELU_BWD, tag="gelu_bwd")
    run(520, 1024, 256, epi=L.EPI_GELU_BWD, quick=True, tag="quickgelu_bwd")
    run(4096, 4096, 4096, tag="4k")
    print(f"correctness phase {time.time() - t0:.1f}s", flush=True)
    T = 65792
    timeit(8192, 8192, 8192, tag="8k^3")
    timeit(T, 3072, 1024, tag="qkv")
    timeit(T, 1024, 1024, epi=L.EPI_RESIDUAL, tag="out_proj+res")
    timeit(T, 4096, 1024, epi=L.EPI_GELU, tag="fc+gelu")
    timeit(T, 1024, 4096, epi=L.EPI_RESIDUAL, tag="proj+res")
    timeit(T, 4096, 1024, epi=L.EPI_GELU_BWD, tag="dgrad proj+gelu'")
    timeit(4096, 1024, T, a_mn=True, b_mn=True, f32=True, split_k=4, tag="wgrad fc")
    timeit(1024, 4096, T, a_mn=True, b_mn=True, f32=True, split_k=4, tag="wgrad proj")
    timeit(3072, 1024, T, a_mn=True, b_mn=True, f32=True, split_k=6, tag="wgrad qkv")
    print("done", flush=True)
