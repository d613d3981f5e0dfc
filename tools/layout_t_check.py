"""Transposed spectrum layout (lp2d_plan_set_spectrum_layout): parity against the natural layout on the same plan, and
time per beam of fft2 / stateful propagate / lens in both layouts."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightprop2d_b200 import Beam2D, um, nm, _lib  # noqa: E402
from lightprop2d_b200 import beam2d as _b  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def rel(a, b):
    return float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt())


def main():
    cases = [(1024, 64, 256), (512, 64, 512), (256, 64, 1024), (512, 128, 256), (256, 128, 512), (2048, 64, 32)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    lib = _lib.load()
    hbm = 6534.8e9
    for n, bits, batch in cases:
        ct = torch.complex64 if bits == 64 else torch.complex128
        B = 8 if bits == 64 else 16
        torch.manual_seed(1)
        f = torch.randn(batch, n, n, dtype=ct, device="cuda")
        _b._PLAN_CACHE.clear()
        b = Beam2D(200 * um * n / 256, n, 632 * nm, init_field=f[:1], complex_bits=bits, use_gpu=True)
        plan = b._plan.handle
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        z = 3e-3
        S = torch.empty_like(f); o = torch.empty_like(f); ST = torch.empty_like(f); oT = torch.empty_like(f)
        I = torch.empty(batch, n, n, dtype=torch.float32 if bits == 64 else torch.float64, device="cuda"); IT = torch.empty_like(I)
        res = {}
        for lay in (0, 1):
            rc = lib.lp2d_plan_set_spectrum_layout(plan, lay)
            if rc != 0:
                print(f"n={n} c{bits}: layout {lay} refused: {lib.lp2d_last_error().decode()}")
                break
            s_, o_, i_ = (S, o, I) if lay == 0 else (ST, oT, IT)
            t_fft = timeit(lambda: _lib.check(lib.lp2d_fft2(plan, f.data_ptr(), s_.data_ptr(), batch, 0, st), "fft2"))
            ff = f.clone()
            _lib.check(lib.lp2d_lens(plan, ff.data_ptr(), s_.data_ptr(), batch, float(b.k0), 0.7, st), "lens")
            lens_s = s_.clone()
            t_lens = timeit(lambda: _lib.check(lib.lp2d_lens(plan, ff.data_ptr(), s_.data_ptr(), batch, float(b.k0), 1e9, st), "lens"))
            _lib.check(lib.lp2d_fft2(plan, f.data_ptr(), s_.data_ptr(), batch, 0, st), "fft2")
            spec0 = s_.clone()
            _lib.check(lib.lp2d_propagate(plan, s_.data_ptr(), o_.data_ptr(), i_.data_ptr(), batch, z, st), "propagate")
            spec1 = s_.clone(); fld = o_.clone(); inten = i_.clone()
            back = torch.empty_like(f)
            _lib.check(lib.lp2d_fft2(plan, spec0.data_ptr(), back.data_ptr(), batch, 1, st), "ifft2")
            t_prop = timeit(lambda: _lib.check(lib.lp2d_propagate(plan, s_.data_ptr(), o_.data_ptr(), None, batch, z, st), "propagate"))
            res[lay] = dict(t_fft=t_fft, t_lens=t_lens, t_prop=t_prop, spec0=spec0, spec1=spec1, fld=fld, inten=inten, lens=lens_s, back=back)
        if len(res) == 2:
            a, t = res[0], res[1]
            tr = lambda x: x.transpose(-1, -2)  # noqa: E731
            tt = torch.empty_like(f)
            _lib.check(lib.lp2d_transpose(_lib.C64 if bits == 64 else _lib.C128, t["spec1"].data_ptr(), tt.data_ptr(), n, batch, st), "transpose")
            print(f"n={n} c{bits} batch={batch}: fft2 {rel(tr(t['spec0']), a['spec0']):.1e}  lens {rel(tr(t['lens']), a['lens']):.1e}  "
                  f"state {rel(tr(t['spec1']), a['spec1']):.1e}  field {rel(t['fld'], a['fld']):.1e}  |E|^2 {rel(t['inten'], a['inten']):.1e}  "
                  f"ifft2(fft2 f) {rel(t['back'], f):.1e} / {rel(a['back'], f):.1e}  lp2d_transpose {rel(tt, tr(t['spec1'])):.1e}")
            P = n * n * B * batch
            for k, mult in (("t_fft", 2), ("t_lens", 3), ("t_prop", 3)):
                print(f"    {k[2:]:5s}: natural {a[k]*1e3/batch:7.2f} us/beam ({mult*P/(a[k]*1e-3)/hbm:.3f})   transposed {t[k]*1e3/batch:7.2f} us/beam ({mult*P/(t[k]*1e-3)/hbm:.3f})", flush=True)
        lib.lp2d_plan_set_spectrum_layout(plan, 0)
        del b


if __name__ == "__main__":
    main()
