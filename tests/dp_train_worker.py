"""Two-rank data-parallel training check, launched by tests/test_gpu_train.py::test_data_parallel_two_gpus through torchrun
(one process per GPU, NCCL): every rank trains on its own shard; after the SUM all-reduce of the flat gradient buffer all
ranks hold the same gradient = the sum of the per-rank gradients, and identical weights after the RMSprop update."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from realtime_style_transfer_b200 import distributed as D, optimizers  # noqa: E402
from test_gpu_train import _dataset, _python_training_model  # noqa: E402


def main():
    rank, world, local = D.env_rank()
    torch.cuda.set_device(local)
    D.init_process_group("nccl", device=torch.device("cuda", local))
    models = _python_training_model()
    # identical initial variables on every rank (rank 0's), as a restored checkpoint would give
    names = sorted(models.inference.weights)
    for k in names:
        t = torch.from_numpy(models.inference.weights[k].copy()).cuda()
        dist.broadcast(t, 0)
        models.inference.set_weights({k: t.cpu().numpy()})
    models.training.compile(optimizer=optimizers.RMSprop())
    batch = _dataset(1, 2, seed=100 + rank)[0]          # a different shard per rank
    tm = models.training
    # local gradient first (no all-reduce): run the native step by hand
    x, y = batch
    tr = tm._get_trainer(2)
    dev = torch.device("cuda", local)
    d = [torch.from_numpy(a).to(dev) for a in (x["content"], x["style"][:, 0], y["content"], y["style"][:, 0])]
    d_l = torch.empty((2, 4), device=dev)
    torch.cuda.synchronize()
    tr.forward_backward(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d_l.data_ptr(), 2)
    local_grad = tm.gradient_tensor().clone()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    expect = torch.stack(gathered).sum(0)
    D.allreduce_sum_(tm.gradient_tensor())
    got = tm.gradient_tensor()
    err = (got - expect).abs().max().item() / max(expect.abs().max().item(), 1e-30)
    assert err < 1e-6, err
    tr.apply_gradients()
    tm._host_stale = True
    tm.sync_to_host()
    # all ranks hold identical trainable variables after the update
    for k in names:
        if k.endswith(("moving_mean", "moving_variance")):
            continue                                    # BatchNorm statistics are per replica (non-synchronised BN)
        t = torch.from_numpy(models.inference.weights[k].copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), k
    # and the public train_step path (which all-reduces internally) keeps them identical
    logs = tm.train_step(batch)
    tm.sync_to_host()
    t = torch.from_numpy(models.inference.weights["residual_block_2/conv1/kernel"].copy()).cuda()
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref)
    dist.barrier()
    if rank == 0:
        print("DP_TRAIN_OK", world, float(np.mean(list(logs.values()))))
    tm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
