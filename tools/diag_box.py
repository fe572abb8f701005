"""Diagnostic: where do box rows differ from the oracle at N = 32768? (scratch helper)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import oracle_lib as ol
from test_oracle_chain import scene
b2s = ge.load_b2s()
n, fs, frames, learn = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 20_000_000, 200, 40
cfg, tones, iq, period = scene(n, fs, frames, learn)
eng = b2s.Engine(0)
band = b2s.Band(eng, cfg)
DENSE = ("psd_db", "noise_sub_db", "avg_db", "box_db")
got = band.push(iq, frames, 1000, period, per_frame=True, dense=DENSE)
ref = ol.OracleChain(cfg).push(iq, frames, 1000, period)
for name in ("avg_db", "box_db"):
    d = np.abs(getattr(got, name) - getattr(ref, name))
    bad = np.argwhere(d > 0.01)
    print(name, "max", d.max(), "n_bad", len(bad))
    if len(bad):
        fr, bins = bad[:, 0], bad[:, 1]
        print(" frames", np.unique(fr)[:20], "bins min/max", bins.min(), bins.max(), "unique bins", np.unique(bins)[:40])
        print(" bins mod 112:", np.unique(bins % 112)[:40], " cta:", np.unique(bins // 112)[:20])
        f0, b0 = bad[0]
        print(" sample got", getattr(got, name)[f0, b0 - 3 : b0 + 4], "ref", getattr(ref, name)[f0, b0 - 3 : b0 + 4])
        print(" avg around", got.avg_db[f0, b0 - 12 : b0 + 13])
