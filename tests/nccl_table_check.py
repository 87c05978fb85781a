"""BASELINE config 4 across ranks on real GPUs: each rank encodes its shard of the patch list, the fused symbol
histogram is all-reduced IN PLACE over NCCL (parallel.allreduce_histogram, the path's only collective), and every
rank must end up with the table one GPU computes over the whole list (get_encoded_distribution.py:113-134).
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_table_check.py
(tests/test_gpu_parity.py::test_nccl_table_allreduce launches it when two GPUs are visible)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from gpu_common import make_codec, patches_from_images  # noqa: E402
from tf_image_compression_b200 import entry, parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    variant = "base_model/reduced_btn_32"
    import tf_image_compression_b200 as T
    from gpu_common import MEAN, STD, params_for
    enc, dec = params_for(variant, "fanin")
    codec = T.Codec(variant, quan_scale=2, mean=MEAN, std=STD, enc_params=enc, dec_params=dec, device=local, compute="tensor")
    patches = patches_from_images(5, 256, 384, 128, seed=200)  # 30 patches, the same list on every rank
    b, e = parallel.shard_range(len(patches))
    prob = entry.get_distribution(codec, patches[b:e])          # encode own shard + NCCL all-reduce of the table
    counts = codec.hist_read()                                  # in place: the device histogram now holds the global table
    # single-GPU answer over the whole list
    codec.hist_reset()
    sym = codec.encode_patches(patches)
    whole = codec.hist_read()
    assert np.array_equal(counts, whole), (rank, counts, whole)
    assert int(counts.sum()) == sym.size and np.allclose(prob, whole / whole.sum())
    # every rank holds the same table
    t = torch.from_numpy(counts.astype(np.int64)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    codec.close()
    dist.barrier()
    if rank == 0:
        print(f"nccl table check ok: world {world}, counts {counts.tolist()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
