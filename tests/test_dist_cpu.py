"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path -- the 8-byte normaliser
all-reduce (ld_head.py:338-341, 362-363 `reduce_mean`) and the log-scalar reduction of `_parse_losses`
(base.py:211-216)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ld_b200 import ops
    from ld_b200.detectors import BaseDetector

    class _TG:   # only .counts is read by loss_normalisers
        counts = torch.tensor([0, 0, 10 + 30 * rank, 0], dtype=torch.int32)
    norm = ops.loss_normalisers(_TG, torch.tensor([2.0 + 4.0 * rank]))
    losses = dict(loss_cls=[torch.tensor(1.0 + rank), torch.tensor(2.0)], loss_bbox=[torch.tensor(0.5 * (rank + 1))],
                  acc=torch.tensor(7.0 * (rank + 1)))
    loss, log = BaseDetector._parse_losses(None, losses)
    q.put((rank, norm.tolist(), float(loss), dict(log)))
    dist.destroy_process_group()


def test_normaliser_and_log_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, norm, loss, log in res:
        # mean over ranks of num_total_pos (10, 40) -> 25; of sum_w + 1e-6 (2, 6) -> 4
        assert abs(norm[0] - 25.0) < 1e-6 and abs(norm[1] - 4.000001) < 1e-5
        # local loss is NOT reduced (it drives the local backward); the log vars are rank means
        assert abs(loss - ((1.0 + rank) + 2.0 + 0.5 * (rank + 1))) < 1e-6
        assert abs(log['loss_cls'] - 3.5) < 1e-6 and abs(log['loss_bbox'] - 0.75) < 1e-6 and abs(log['acc'] - 10.5) < 1e-6
        assert abs(log['loss'] - 4.25) < 1e-6
