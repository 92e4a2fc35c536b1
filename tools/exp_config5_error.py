"""Where does the mask error of the config 5 fixture come from?  (GPU box; the oracle runs in fp32 on the same GPU with TF32 off.)

Four (encoder, decoder) combinations against (oracle encoder, oracle decoder):
  ours/ours   the product path
  ref/ours    decoder error alone
  ours/ref    encoder error alone, propagated through an exact decoder
and the density of reference logits near zero, which is what converts a logit error into an IoU loss."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

from oracle import sam_oracle as O
from ullsam_b200 import synth
from ullsam_b200.engine import Engine

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mllm.npz"))
seed, P = int(g["seed"]), int(g["P"])
cfg = O.config("vit_b")
sd = synth.make_state_dict(cfg, seed)
sdc = {k: v.cuda() for k, v in sd.items()}
images, _ = synth.make_images(1, 1024, seed=seed + 11)
pts, lab = synth.make_points(P, 1024, seed=seed + 12)
dense_map = torch.from_numpy(g["dense_map_f16"]).float().cuda()
eng = Engine(model="vit_b")
eng.load_state_dict(sd)


def iou(a, b):
    a, b = a > 0, b > 0
    return (a & b).sum().item() / max((a | b).sum().item(), 1)


with torch.no_grad():
    emb_ref = O.encoder_forward(sdc, images.cuda(), cfg)
    emb = eng.encoder_forward(images.cuda())
    print("embedding rel-L2 %.3e" % ((emb.double() - emb_ref.double()).norm() / emb_ref.double().norm()).item())
    pe = O.dense_pe(sdc, cfg)
    sparse = O.embed_points(sdc, pts.cuda(), lab.cuda(), cfg)
    dense = O.dense_from_llm(sdc, dense_map).repeat(P, 1, 1, 1)
    for mm in (True, False):
        ref, _ = O.decoder_forward(sdc, emb_ref, pe, sparse, dense, mm, cfg)
        a, _ = eng.decoder_forward(emb, pe, sparse, dense, mm)
        b, _ = eng.decoder_forward(emb_ref, pe, sparse, dense, mm)
        c, _ = O.decoder_forward(sdc, emb, pe, sparse, dense, mm, cfg)
        sc = ref.abs().mean().item()
        print(f"multimask={mm}: |ref logit| mean {sc:.3e} max {ref.abs().max().item():.3e}")
        for tag, x in (("ours/ours", a), ("ref/ours ", b), ("ours/ref ", c)):
            e = (x - ref)
            print(f"  {tag}: IoU {iou(x, ref):.4f}  logit err rms {e.pow(2).mean().sqrt().item():.3e} max {e.abs().max().item():.3e}"
                  f"  (rms / mean|logit| {e.pow(2).mean().sqrt().item() / sc:.3e})")
        for t in (5e-4, 1e-3, 2e-3, 4e-3):
            print(f"    fraction of reference logits with |x| < {t:g}: {(ref.abs() < t).float().mean().item():.4f}")
eng.close()
