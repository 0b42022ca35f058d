"""Multi-GPU path on real GPUs (skipped on a 1-GPU box): NCCL all-gather of the packed partials + merge kernel must equal
the oracle's single scan; frame-sharded extraction must equal the oracle frame by frame."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %r)
from orb_slam2_refactored_b200 import api, synth, distributed as D
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
q, t = synth.planted_descriptors(33, 5000, 240007, dup_every=2)
b, e = D.shard_range(len(t), world, rank)
dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t[b:e]).cuda()
idx, best, second, match = D.knn2_sharded(dq, dt, b)
torch.cuda.synchronize()
frames = np.stack([synth.image(100 + s, 640, 480) for s in range(2 * world)])
ex = api.ORBextractor(nfeatures=1000, device=lr)
fb, fe, kps, desc = D.extract_sharded(ex, frames)
if rank == 0:
    np.savez(sys.argv[1], idx=idx.cpu().numpy(), best=best.cpu().numpy().view(np.uint16), second=second.cpu().numpy().view(np.uint16),
             match=match.cpu().numpy())
np.savez(sys.argv[1] + '.frames%%d.npz' %% rank, fb=fb, fe=fe, **{'k%%d' %% i: k.view(np.uint8) for i, k in enumerate(kps)},
         **{'d%%d' %% i: d for i, d in enumerate(desc)})
dist.barrier()
dist.destroy_process_group()
''' % ROOT


def test_two_rank_knn_and_frame_sharding(tmp_path, orbx, oracle_port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    from orb_slam2_refactored_b200 import synth
    world = 2
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    out = str(tmp_path / 'out.npz')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
                        '--master-port', '29631', str(script), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    q, t = synth.planted_descriptors(33, 5000, 240007, dup_every=2)
    want = oracle_port.knn2(q, t, 50, 0.6, threads=8)
    for name, w in zip(('idx', 'best', 'second', 'match'), want):
        assert np.array_equal(got[name], w), name
    e = oracle_port.extractor(1000)
    for rank in range(world):
        fr = np.load(out + f'.frames{rank}.npz')
        for i, fidx in enumerate(range(int(fr['fb']), int(fr['fe']))):
            okps, odesc = e.extract(synth.image(100 + fidx, 640, 480))
            assert fr[f'k{i}'].tobytes() == okps.tobytes() and np.array_equal(fr[f'd{i}'], odesc)
