"""Decoder alone (32 prompts, box prompts) a few times: for timeline builds and quick timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sam_oracle as O
from ullsam_b200 import synth
from ullsam_b200.engine import Engine
cfg = O.config("vit_b")
sd = synth.make_state_dict(cfg, 1234)
eng = Engine(model="vit_b"); eng.load_state_dict(sd)
B = 32
emb = torch.randn(B, 256, 64, 64, device="cuda") * 0.1
sdc = {k: v.cuda() for k, v in sd.items() if k.startswith("prompt_encoder.")}
pe = O.dense_pe(sdc, cfg)
_, boxes = synth.make_images(B, 1024, seed=3)
sparse = O.embed_boxes(sdc, boxes.cuda(), cfg)
dense = O.dense_no_mask(sdc, B, cfg).contiguous()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    m, s = eng.decoder_forward(emb, pe, sparse, dense, False)
torch.cuda.synchronize()
if n > 1:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.decoder_forward(emb, pe, sparse, dense, False)
    e1.record(); torch.cuda.synchronize()
    print(f"decoder, {B} prompts: {e0.elapsed_time(e1) / 10:.3f} ms")
eng.close()
