import ctypes as C, os, sys, time, numpy as np
ROOT=os.environ.get("GRAFT_REPO_ROOT","/root/repo"); sys.path.insert(0,ROOT)
import torch
from ssrlcv_b200 import types as T
from ssrlcv_b200.match_factory import Image, MatchFactory
from ssrlcv_b200.unity import Unity, cpu, gpu
from tests.golden_data import load_everest
ev=load_everest(); q,t,s=ev.views2[0],ev.views2[1],ev.seed2
F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
for engine in (0,1):
    mf=MatchFactory(ev.rel_threshold, ev.abs_threshold, engine=engine)
    qu,tu=Unity(q,T.FEATURE,cpu),Unity(t,T.FEATURE,cpu); qu.setMemoryState(gpu); tu.setMemoryState(gpu)
    qi,ti=Image(ev.img_ids[0],ev.cameras[0]),Image(ev.img_ids[1],ev.cameras[1])
    def timed(name,fn):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(5): r=fn()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5*1e3
        print(f"engine {engine} {name:34s} wall {dt:7.3f} ms records {r.size()}")
    timed("DMatch double unseeded", lambda: mf.generateDistanceMatchesDoubleConstrained(qi,qu,ti,tu,25.0,5.0))
    timed("DMatch F unseeded", lambda: mf.generateDistanceMatchesConstrained(qi,qu,ti,tu,25.0,F))
    timed("DMatch unmasked unseeded", lambda: mf.generateDistanceMatches(qi,qu,ti,tu))
