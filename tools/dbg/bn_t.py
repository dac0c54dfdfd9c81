import ctypes, sys, torch
sys.path.insert(0, ".")
import gd_mae_b200
from gd_mae_b200 import _lib as L
lib = L.lib()
torch.manual_seed(0)
N, C = 3001, 128
u = torch.randn(N, C, device="cuda")
gamma = torch.rand(C, device="cuda") + 0.5
beta = torch.rand(C, device="cuda") - 0.5
count = float(3 * N)
ws = torch.empty(lib.gdmae_batchnorm_workspace_bytes(C), dtype=torch.uint8, device="cuda")
def fwd_old():
    out = torch.empty_like(u); mean = torch.empty(C, device="cuda"); rstd = torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_fwd(L.P(u), L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(count), L.f32(1e-3), L.f32(0.01), 1, L.P(out), L.P(mean), L.P(rstd), None, None, L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "f")
    return out, mean, rstd
def fwd_new(yd, od):
    y = u.to(yd).contiguous()
    out = torch.empty((N, C), dtype=od, device="cuda"); mean = torch.empty(C, device="cuda"); rstd = torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_fwd_t(L.P(y), int(yd == torch.bfloat16), L.P(gamma), L.P(beta), L.i64(N), C, ctypes.c_double(count), L.f32(1e-3), L.f32(0.01), 1, L.P(out), int(od == torch.bfloat16), L.P(mean), L.P(rstd), None, None, L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "f")
    return out, mean, rstd
o0, m0, r0 = fwd_old()
for yd in (torch.float32, torch.bfloat16):
    for od in (torch.float32, torch.bfloat16):
        o1, m1, r1 = fwd_new(yd, od)
        print("fwd", yd, od, (o1.float() - o0).abs().max().item(), (m1 - m0).abs().max().item(), (r1 - r0).abs().max().item())
dout = torch.randn(N, C, device="cuda").bfloat16()
def bwd_old():
    dy = torch.empty_like(u); dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_bwd(L.P(u), L.P(beta), L.P(dout.float().contiguous()), L.P(gamma), L.P(m0), L.P(r0), L.i64(N), C, ctypes.c_double(count), 1, None, None, L.P(dy), None, L.P(dg), L.P(db), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "b")
    return dy, dg, db
def bwd_new(yd, dd, gd):
    y = u.to(yd).contiguous(); d = dout.to(dd).contiguous()
    dy = torch.empty((N, C), dtype=gd, device="cuda"); dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
    L.check(lib.gdmae_batchnorm_relu_bwd_t(L.P(y), int(yd == torch.bfloat16), L.P(beta), L.P(d), int(dd == torch.bfloat16), L.P(gamma), L.P(m0), L.P(r0), L.i64(N), C, ctypes.c_double(count), 1, None, None, L.P(dy), int(gd == torch.bfloat16), L.P(dg), L.P(db), L.P(ws), ctypes.c_size_t(ws.numel()), L.stream()), "b")
    return dy, dg, db
y0, g0, b0 = bwd_old()
for yd in (torch.float32, torch.bfloat16):
    for dd in (torch.float32, torch.bfloat16):
        for gd in (torch.float32, torch.bfloat16):
            y1, g1, b1 = bwd_new(yd, dd, gd)
            print("bwd", yd, dd, gd, (y1.float() - y0).abs().max().item(), (g1 - g0).abs().max().item(), (b1 - b0).abs().max().item(), y0.abs().max().item(), g0.abs().max().item())
