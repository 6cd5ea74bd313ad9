import sys, hashlib
sys.path[:0] = ["/root/repo/2d-vq-ae-2_b200", "/root/repo/profiles"]
import torch
import vqae_b200
from vqae_b200 import synthetic as S
from vqae_b200.extract import encode_patches
from slide_bench import device_patches
from vqae_b200 import engine as E
if len(sys.argv) > 2 and sys.argv[2] == "chain":
    E.TRUNK_RESIDENT = False
dev = torch.device("cuda:0")
model = vqae_b200.build_vqae(n_down=3).eval()
model.load_state_dict(S.make_state_dict(model.state_dict(), seed=1, regime="perturbed"))
m = vqae_b200.set_precision(model.to(dev), sys.argv[1] if len(sys.argv) > 1 else "fp16")
enc = m.encoder
NB = 24
batches = [device_patches(256 * k, 256, dev) for k in range(NB)]
torch.cuda.synchronize()
print("inputs", hashlib.sha256(torch.cat(batches).cpu().numpy().tobytes()).hexdigest()[:12])
with torch.no_grad():
    ref = []
    for b in batches:                       # one by one, synchronised
        ref.append(encode_patches(enc, b).clone())
        torch.cuda.synchronize()
    for trial in range(6):
        outs = [encode_patches(enc, b) for b in batches]      # back to back
        torch.cuda.synchronize()
        bad = [k for k in range(NB) if not torch.equal(outs[k], ref[k])]
        print("trial", trial, "batches differing from the synchronised run:", bad,
              [int((outs[k] != ref[k]).sum()) for k in bad])
