"""Layer-by-layer check of the image-encoder conv kernels against torch conv2d (fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodalfilter_b200 import ops
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37
img = torch.rand(n, 32, 32, device=dev) * 2 - 1
c1 = torch.nn.Conv2d(1, 32, 5, padding=2).to(dev)
c2a = torch.nn.Conv2d(32, 32, 3, padding=1).to(dev)
c2b = torch.nn.Conv2d(32, 32, 3, padding=1).to(dev)
c3 = torch.nn.Conv2d(32, 16, 3, padding=1).to(dev)
c4 = torch.nn.Conv2d(16, 8, 3, padding=1).to(dev)

def unmap(m, ch):
    """planes -> (n, ch, 32, 32) fp32 (hi + lo)"""
    per = ops.enc_map_bytes(ch)
    t = m.view(n, per).view(torch.bfloat16).reshape(n, ch // 8, 2, 1280, 8).float()
    v = (t[:, :, 0] + t[:, :, 1])[:, :, 64:64 + 1088].reshape(n, ch // 8, 32, 34, 8)[:, :, :, :32]
    return v.permute(0, 1, 4, 2, 3).reshape(n, ch, 32, 32)

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())

with torch.no_grad():
    x_ref = F.relu(c1(img[:, None]))
    mx = ops.enc_new_map(n, 32, dev)
    ops.enc_stem(img.contiguous(), ops.enc_pack_stem(c1), mx)
    torch.cuda.synchronize()
    print("stem rel err", rel(unmap(mx, 32), x_ref))
    t_ref = F.relu(c2a(x_ref))
    mt = ops.enc_new_map(n, 32, dev)
    ops.enc_conv3x3(n, 32, 32, mx, ops.enc_pack_conv3x3(c2a), relu=True, out_map=mt)
    torch.cuda.synchronize()
    print("conv 32->32 rel err", rel(unmap(mt, 32), t_ref))
    y_ref = F.relu(c2b(t_ref) + x_ref)
    my = ops.enc_new_map(n, 32, dev)
    ops.enc_conv3x3(n, 32, 32, mt, ops.enc_pack_conv3x3(c2b), res_map=mx, relu=True, out_map=my)
    torch.cuda.synchronize()
    print("conv 32->32 + residual rel err", rel(unmap(my, 32), y_ref))
    z_ref = F.relu(c3(y_ref))
    mz = ops.enc_new_map(n, 16, dev)
    ops.enc_conv3x3(n, 32, 16, my, ops.enc_pack_conv3x3(c3), relu=True, out_map=mz)
    torch.cuda.synchronize()
    print("conv 32->16 rel err", rel(unmap(mz, 16), z_ref))
    o_ref = c4(z_ref)
    out = torch.empty(n, 8, 32, 32, device=dev)
    ops.enc_conv3x3(n, 16, 8, mz, ops.enc_pack_conv3x3(c4), relu=False, out_nchw=out)
    torch.cuda.synchronize()
    print("conv 16->8 (nchw) rel err", rel(out, o_ref))

if len(sys.argv) > 2:
    n = int(sys.argv[2])
    img = torch.rand(n, 32, 32, device=dev) * 2 - 1
    mx, mt, my = (ops.enc_new_map(n, 32, dev) for _ in range(3)); mz = ops.enc_new_map(n, 16, dev)
    out = torch.empty(n, 8, 32, 32, device=dev)
    ws, w2a, w2b, w3, w4 = ops.enc_pack_stem(c1), ops.enc_pack_conv3x3(c2a), ops.enc_pack_conv3x3(c2b), ops.enc_pack_conv3x3(c3), ops.enc_pack_conv3x3(c4)
    def run():
        ops.enc_stem(img, ws, mx)
        ops.enc_conv3x3(n, 32, 32, mx, w2a, relu=True, out_map=mt)
        ops.enc_conv3x3(n, 32, 32, mt, w2b, res_map=mx, relu=True, out_map=my)
        ops.enc_conv3x3(n, 32, 16, my, w3, relu=True, out_map=mz)
        ops.enc_conv3x3(n, 16, 8, mz, w4, relu=False, out_nchw=out)
    run(); torch.cuda.synchronize()
    ops.PROFILE.reset(enabled=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    prof = ops.PROFILE.collect()
    print(f"{n} images: {e0.elapsed_time(e1)/2:.2f} ms per pass", {k: round(v['avg_ms'], 3) for k, v in prof['kernels'].items()})
    seq = torch.nn.Sequential(c1, torch.nn.ReLU(), c2a, torch.nn.ReLU(), c2b, torch.nn.ReLU(), c3, torch.nn.ReLU(), c4)
    with torch.no_grad():
        seq(img[:, None]); torch.cuda.synchronize(); e0.record(); seq(img[:, None]); e1.record(); torch.cuda.synchronize()
    print(f"cuDNN fp32 same layers (no residual add): {e0.elapsed_time(e1):.2f} ms")

# fused trunk
n = int(sys.argv[3]) if len(sys.argv) > 3 else 300
img = torch.rand(n, 32, 32, device=dev) * 2 - 1
with torch.no_grad():
    ref = c4(F.relu(c3(F.relu(c2b(F.relu(c2a(F.relu(c1(img[:, None]))))) + F.relu(c1(img[:, None]))))))
w = ops.enc_pack_trunk([c1, c2a, c2b, c3, c4]); sc = ops.enc_trunk_scratch(dev)
out = ops.enc_trunk(img, w, sc, 8); torch.cuda.synchronize()
print("fused trunk rel err", rel(out, ref), "n", n)
out2 = ops.enc_trunk(img, w, sc, 8); torch.cuda.synchronize()
print("second run identical", bool(torch.equal(out, out2)))
if len(sys.argv) > 2:
    n = int(sys.argv[2]); img = torch.rand(n, 32, 32, device=dev) * 2 - 1
    ops.enc_trunk(img, w, sc, 8); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ops.enc_trunk(img, w, sc, 8); ops.enc_trunk(img, w, sc, 8); e1.record(); torch.cuda.synchronize()
    print(f"fused trunk, {n} images: {e0.elapsed_time(e1)/2:.2f} ms")
