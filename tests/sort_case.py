import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "crowd-coachable-recommendations_b200"))
import torch
from ccr_b200 import engine
S = torch.randn((64, 1048576), generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
engine.argsort_scores(S, None); torch.cuda.synchronize()
engine.argsort_scores(S, None); torch.cuda.synchronize()
