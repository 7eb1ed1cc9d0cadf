"""Debug (GPU box): which host call stalls when a step's enqueue time spikes?  Profiles every step with cProfile and prints the
top cumulative entries of the slow ones."""
import cProfile, gc, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from odwscl_b200.config import get_cfg_defaults
from odwscl_b200.modeling import build_detection_model
from odwscl_b200.structures import BoxList
from odwscl_b200.synth import synth_batch
torch.manual_seed(0)
model = build_detection_model(get_cfg_defaults()).cuda().train()
opt = bench.make_optimizer(model)
images, rois, boxes, labels = synth_batch(2, 2000, 1000, 600, 21, seed=1234, pin=True)
targets = []
for lab in labels:
    t = BoxList(torch.zeros((len(lab), 4)), (1000, 600), "xyxy"); t.add_field("labels", torch.as_tensor(lab)); targets.append(t)
images_d, rois_d = images.cuda(), rois.cuda()
props = [BoxList(r[:, 1:], (1000, 600), "xyxy") for r in rois_d.split([2000, 2000])]
ev = model.roi_heads.loss_evaluator
ev.speculative_k = True; ev.k_margin, ev.k_granule = 1.5, 128
def step():
    losses, _ = model(images_d, targets, props)
    total = sum(losses.values())
    opt.zero_grad(set_to_none=True)
    total.backward()
    if ev.overflow is not None:
        opt.found_inf = ev.overflow
    opt.step()
for _ in range(12):
    step()
torch.cuda.synchronize()
gc.collect(); gc.freeze()
slow = 0
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable(); step(); pr.disable()
    dt = (time.perf_counter() - t0) * 1e3
    if dt > 30:
        slow += 1
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(8)
        print("== step %d host %.1f ms, K cap %s, gc counts %s" % (i, dt, ev._k_cap, gc.get_count()))
        print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:2500])
torch.cuda.synchronize()
print("slow steps:", slow)
