"""Host-side cost of enqueueing one forward (python + ctypes + graph launch), vs GPU time per step."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from bench import WORKLOADS, make_inputs
from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict
kwargs, batch, desc, kind = WORKLOADS["B64"]
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd = make_state_dict(PathConfig(**kwargs), 0)
models = []
for _ in range(S):
    m = MCM(**kwargs, share_sm=S > 1); m.load_state_dict(sd); m.cuda().eval(); models.append(m)
imgs, scores = make_inputs(kwargs, batch, 0, 2)
imgs = [t.cuda() for t in imgs]; scores = [t.cuda() for t in scores]
side = [torch.cuda.Stream() for _ in range(S)]
def step(i):
    with torch.cuda.stream(side[i % S]):
        return models[i % S](imgs[i % 2], scores[i % 2])
for i in range(6 * S): step(i)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for i in range(n): out = step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"S={S}: host enqueue {1e3*(t1-t0)/n:.3f} ms/step, total {1e3*(t2-t0)/n:.3f} ms/step -> {batch*n/(t2-t0):.0f} images/s")
# raw C call only (no python output handling): reuse the last outputs through forward_host-less path is not exposed; time the ctypes call
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(50): out = step(i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
