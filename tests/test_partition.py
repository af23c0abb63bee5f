"""Multi-GPU host logic on CPU: sample-block partition + the single film reduce, world size 2 over gloo.
The per-rank "render" runs the csrc sources in the host simulation (tests/hostsim), which shares the sample ->
PCG stream mapping with the CUDA build, so the reduced film must equal the single-process render."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lajolla_public_b200 import partition


def test_sample_ranges_cover_without_overlap():
    for total in (0, 1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            r = partition.sample_ranges(total, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    assert partition.weak_range(3, 8, 128) == (1024, 384, 512)
    with pytest.raises(ValueError):
        partition.weak_range(8, 8, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ljs_path, total_spp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import hostsim_lib
    import lajolla_public_b200 as lj
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with hostsim_lib.simulated():
            sc = lj.parse_scene(ljs_path)
            film = torch.zeros((sc.height, sc.width, 3), dtype=torch.float32)  # hostsim "device" memory is host memory
            st = partition.render_partitioned(sc, total_spp, rank, world, film, dist=dist, pool_paths=1 << 15)
            begin, end = partition.sample_ranges(total_spp, world)[rank]
            assert st.samples == sc.width * sc.height * (end - begin)
        if rank == 0:
            np.save(os.path.join(out_dir, "reduced.npy"), film.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_render_equals_single(oracle, tmp_path):
    import hostsim_lib
    import lajolla_public_b200 as lj
    ljs_path = oracle.scene_ljs("cbox")
    total_spp = 3  # uneven split: rank 0 renders 2 samples, rank 1 one
    mp.spawn(_worker, args=(2, _free_port(), ljs_path, total_spp, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    with hostsim_lib.simulated():
        single = lj.parse_scene(ljs_path).render(spp=total_spp, pool_paths=1 << 15)
    assert np.allclose(reduced, single, rtol=1e-4, atol=1e-5)
