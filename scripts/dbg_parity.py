"""Debug helper (GPU box): strict-mode cfg1 product vs oracle -- tau, close-set sizes, augmented embeddings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import test_gpu_parity_configs as T
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_voc07_bs2_2000"
mode = sys.argv[2] if len(sys.argv) > 2 else "strict"
_, _, boxes, labels, ref_losses, tr = T._oracle(name)
import odwscl_b200.modeling.vgg16 as V
_orig_neck = V.VGG16FC67ROIFeatureExtractor.forward_neck
AUG = {}
def _neck(self, x):
    AUG["x"] = x.detach().clone()
    AUG["rows"] = self._aug_rows.detach().clone() if hasattr(self, "_aug_rows") else None
    return _orig_neck(self, x)
V.VGG16FC67ROIFeatureExtractor.forward_neck = _neck
got, cap, ev = T._run_product(name, mode)
st = ev.last_state
P = st.P
offA = st.offA.cpu().numpy(); K = int(offA[P])
pi, pc = st.pair_img.cpu().numpy(), st.pair_cls.cpu().numpy()
tau = st.tau.cpu().numpy(); amax = st.amax.cpu().numpy()
E = st.E.cpu()
print("K", K, "P", P, "pairs", list(zip(pi.tolist(), pc.tolist())))
for p in range(P):
    b, c = int(pi[p]), int(pc[p])
    I = tr["trace"]["phaseA_idx"][(b, c)]
    print("pair", p, (b, c), "cntA", offA[p + 1] - offA[p], "oracle", len(I))
    for kind, base in (("drop", 0), ("noise", K)):
        eo = tr["emb"][(b, c, kind)]
        ep = E[base + offA[p]:base + offA[p + 1]]
        print("   ", kind, "max|dE|", float((eo - ep).abs().max()) if eo.shape == ep.shape else ("shape", eo.shape, ep.shape))
    for i in range(3):
        key = (b, i, c)
        print("    i", i, "amax", amax[p, i], tr["trace"]["argmax"][key], "tau", tau[p, i], tr["trace"]["tau"][key],
              "close_o", len(tr["trace"]["close"][key]), "inst_o", len(tr["inst"][b][i][c]), "inst_p", int(st.inst_cnt[p, i]))
print(got); print(ref_losses)
# ---- where do the augmented embeddings diverge?  product's aug INPUT vs a CPU recomputation from the oracle's pooled
from tests.helpers import KeyedSource
from oracle import oracle as orc
ks = KeyedSource(99); ks.dropblock_centres(sum(b.shape[0] for b in boxes), 3)
rows = st.rowsA[:K].cpu()
Xo = tr["pooled"][rows]
print("pooled rows product vs oracle", float((cap["pooled"].cpu()[rows] - Xo).abs().max()))
for p in range(P):
    r = rows[offA[p]:offA[p + 1]]
    d = orc.dropblock(tr["pooled"][r], ks.dropblock_centres_rows(r, 1), 1)
    e = tr["emb"][(int(pi[p]), int(pc[p]), "drop")]
    Fo = tr["simf"][r]
    ep = E[offA[p]:offA[p + 1]]
    cos = lambda a, b: float((a * b).sum(1).mean())
    print("pair", p, "cos(F, E_drop) oracle", cos(Fo, e), "product", cos(Fo, ep), "cos(F,E_noise) oracle",
          cos(Fo, tr["emb"][(int(pi[p]), int(pc[p]), "noise")]), "product", cos(Fo, E[K + offA[p]:K + offA[p + 1]]))

aug = AUG["x"].cpu(); Kp = aug.shape[0] // 2
print("aug shape", aug.shape, "K", K, "rows equal", torch.equal(AUG["rows"].cpu()[:K], rows))
dp = aug[:K]; npart = aug[Kp:Kp + K]
print("product drop part: frac zero bins", float((dp.abs().sum(1) == 0).float().mean()), "noise part ratio stats",
      float((npart / cap["pooled"].cpu()[rows].clamp(min=1e-6)).std()))
for p in range(P):
    r = rows[offA[p]:offA[p + 1]]
    d = orc.dropblock(tr["pooled"][r], ks.dropblock_centres_rows(r, 1), 1)
    nz = ks.noise_rows(r, d.shape) * tr["pooled"][r] + tr["pooled"][r]
    print("pair", p, "max|aug_drop - cpu|", float((dp[offA[p]:offA[p + 1]] - d).abs().max()), "max|aug_noise - cpu|",
          float((npart[offA[p]:offA[p + 1]] - nz).abs().max()), "scale", float(d.abs().max()))
