import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import ops, _lib
from music2dance_b200.ops import Mat
from music2dance_b200.nets import ConvLayer
DEV = "cuda:0"
def cl(x): return x.permute(0, 2, 1).contiguous().to(DEV)
def ncl(m, B, L, C): return m.t.view(B, L, C).permute(0, 2, 1).cpu()
def err(a, b): return float((a - b).abs().max() / b.abs().max())
for case in [(32, 64, 25, 4, 11, 19200, 7), (64, 128, 25, 4, 11, 4800, 16), (64, 128, 25, 4, 11, 4800, 32), (128, 100, 7, 1, 3, 1000, 24)]:
    Cin, Cout, k, s, p, L, B = case
    g = torch.Generator().manual_seed(0)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    wd, bd = w.to(DEV), b.to(DEV)
    lay = ConvLayer("t", wd, bd, torch.zeros_like(wd), torch.zeros_like(bd), Cin, Cout, k, s, p, L)
    lay.pack()
    x = torch.randn(B, Cin, L, generator=g); v = torch.randn(B, Cin, L, generator=g)
    y_ref = F.relu(F.conv1d(x, w, b, stride=s, padding=p)); Lout = y_ref.shape[-1]
    t_ref = F.conv1d(v, w, None, stride=s, padding=p) * (y_ref > 0).float()
    scratch = torch.empty(1 << 22, device=DEV)
    X, V = Mat.of(cl(x), B, L, Cin), Mat.of(cl(v), B, L, Cin)
    for tag in ("inplace", "separate", "inplace-again"):
        Y = Mat.of(torch.empty(B, Lout, Cout, device=DEV), B, Lout, Cout)
        Y2 = Mat.of(torch.empty(B, Lout, Cout, device=DEV), B, Lout, Cout)
        n0 = _lib.load().m2d_halo_persist_launch_count()
        lay.fwd(X, Y, act=1, ws=scratch, y2=Y2)
        torch.cuda.synchronize()
        e_f = err(ncl(Y2, B, Lout, Cout), y_ref)
        if tag == "separate":
            M = Mat.of(Y2.t.clone(), B, Lout, Cout)
            lay.fwd(V, Y2, bias=False, ws=scratch, mask=M, mask_mode=1)
        else:
            lay.fwd(V, Y2, bias=False, ws=scratch, mask=Y2, mask_mode=1)
        torch.cuda.synchronize()
        n1 = _lib.load().m2d_halo_persist_launch_count()
        got = ncl(Y2, B, Lout, Cout)
        d = (got - t_ref).abs()
        bad = (d > 1e-3 * float(t_ref.abs().max())).nonzero()
        print(case, tag, "persist launches", n1 - n0, "fwd err %.2e tangent err %.2e" % (e_f, err(got, t_ref)), "bad", bad.shape[0],
              "first", bad[:4].tolist(), "last", bad[-2:].tolist(), flush=True)
