"""Multi-GPU render inside the library (lj_init(ids, n) + lj_render(num_gpus, split, reduce)): needs >= 2 GPUs in one
process; skipped otherwise (run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

import lajolla_public_b200 as lj
from lajolla_public_b200 import abi

pytestmark = pytest.mark.gpu

LJ_SPLIT_SPP, LJ_SPLIT_TILES = 1, 2
LJ_REDUCE_P2P, LJ_REDUCE_NCCL = 1, 2


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name,spp", [("cbox", 16), ("volpath_test6", 8)])
def test_library_multi_gpu_matches_single(oracle, name, spp):
    """The 2-GPU film (spp split and tile split, P2P-fused and NCCL reductions) equals the 1-GPU film: the same
    (pixel, sample) streams are rendered, only on different devices, so the images agree to fp32 summation order."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    from lajolla_public_b200 import ljs
    desc = ljs.load(oracle.scene_ljs(name))
    sc = lj.Scene(desc, device=[0, 1])
    one = sc.render(spp=spp, num_gpus=1)
    s1 = sc.last_stats
    assert s1.gpus_used == 1
    scale = np.abs(one).mean()
    for split in (LJ_SPLIT_SPP, LJ_SPLIT_TILES):
        for reduce in (LJ_REDUCE_P2P, LJ_REDUCE_NCCL):
            two, var = sc.render(spp=spp, num_gpus=2, split=split, reduce=reduce, variance=True)
            st = sc.last_stats
            assert st.gpus_used == 2 and st.samples == s1.samples, (split, reduce, st.samples, s1.samples)
            assert np.all(np.isfinite(two))
            err = np.abs(two - one).max()
            if name == "cbox":
                assert err <= 1e-4 * max(np.abs(one).max(), 1.0), (split, reduce, err)
            else:  # the volpath NEE walk seeds are per path as well: same bound
                assert err <= 1e-4 * max(np.abs(one).max(), 1.0), (split, reduce, err)
            assert np.all(var >= 0) and np.isfinite(var).all()
    assert scale > 0
    sc.close()


def test_num_gpus_beyond_init_is_rejected(oracle):
    from lajolla_public_b200 import ljs
    sc = lj.Scene(ljs.load(oracle.scene_ljs("cbox")), device=0)
    with pytest.raises(lj.LajollaError):
        sc.render(spp=1, num_gpus=2)
    sc.close()
