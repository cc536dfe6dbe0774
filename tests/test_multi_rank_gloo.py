"""N > 1 host logic on CPU: world_size-2 gloo run of the sample partition + single film reduce
(jet-pbrt_b200/multi_gpu.py), with the CPU oracle standing in for the per-rank renderer."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as ge

    pkg = ge.load_package()
    orc = ge.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "jet-pbrt_b200"))
    import multi_gpu

    sc = pkg.HostScene.builtin("cornell", 24, 24)
    ps = orc.Oracle("port").scene(sc)

    def render_fn(begin, count):
        film, _ = ps.render_counter(begin, count, 5, numthreads=2)
        return torch.from_numpy(film.copy())

    film, spp_total = multi_gpu.render_partitioned(render_fn, 3 * world, dist)
    if rank == 0:
        np.save(Path(out_dir) / "reduced.npy", film.numpy())
        np.save(Path(out_dir) / "final.npy", multi_gpu.finalize(film.clone(), spp_total).numpy())
        assert spp_total == 3 * world
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_rank_render(tmp_path, pkg, port):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    final = np.load(tmp_path / "final.npy")
    sc = pkg.HostScene.builtin("cornell", 24, 24)
    whole, _ = port.scene(sc).render_counter(0, 6, 5, numthreads=4)
    # N-rank reduced film == the 6-sample film up to float reassociation (SURVEY.md 8e "Validation")
    np.testing.assert_allclose(reduced, whole, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(final, np.clip(whole / 6, 0, 1), rtol=1e-5, atol=1e-6)
    assert final.max() <= 1.0 and (reduced.max() > 1.0)  # the clamp happens after the reduce, on the mean


def test_sample_partitions(pkg):
    """jpbrt_sample_partition (the library's own split) and its Python restatement: contiguous, complete, even."""
    sys.path.insert(0, str(ROOT / "jet-pbrt_b200"))
    import multi_gpu

    for total, world in [(50, 8), (4096, 8), (7, 3), (3, 4), (50, 1), (0, 2)]:
        parts = [pkg.sample_partition(total, r, world) for r in range(world)]
        assert parts == [multi_gpu.split_samples(total, r, world) for r in range(world)]
        assert sum(c for _, c in parts) == total and max(c for _, c in parts) - min(c for _, c in parts) <= 1
        assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def test_comm_entry_points_fail_cleanly_without_a_gpu(pkg):
    """The NCCL layer of the ABI: bad arguments are rejected before anything touches a device."""
    C = pkg.C
    assert pkg.lib.jpbrt_comm_unique_id(None, 128) == -1
    small = C.create_string_buffer(8)
    assert pkg.lib.jpbrt_comm_unique_id(small, 8) == -1
    assert pkg.lib.jpbrt_comm_init(None, small, 128, 0, 2) == -1
    assert pkg.lib.jpbrt_comm_rank(None) == 0 and pkg.lib.jpbrt_comm_size(None) == 1
    assert pkg.lib.jpbrt_reduce_film(None) == -1
    if pkg.device_count() == 0:
        sc = pkg.HostScene.builtin("cornell", 8, 8)
        out = np.zeros((8, 8, 3), np.float32)
        rc = pkg.lib.jpbrt_render_multi(sc.desc, 0, 1, 1, 2, out.ctypes.data_as(C.POINTER(C.c_float)), None, None)
        assert rc in (-1, -2, -3)  # 2 GPUs requested, none visible: an error code, never a CPU render
        assert not out.any()
