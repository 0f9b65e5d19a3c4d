import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch, ppad_b200
from ppad_b200 import layout
env = ppad_b200.BatchedPAD(1 << 20, 5, 6, seed=1); env.reset(step=0)
pipe = env.host_pipe(slab_orbs=54, compact=True, slab_format='chunks6')
orbs = np.random.default_rng(0).integers(0, 6, size=(env.n, 54)).astype(np.uint8)
pipe.slab.copy_(torch.from_numpy(layout.pack_slab_chunks6(orbs).view(np.int16)))
pipe.step(eval_moves=True, step=1)
print(pipe.learn_dictionary(max_entries=4096))
pipe.step(eval_moves=True, step=2)
res = pipe.results()
print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in res.items()}, res['reward'].sum())
