"""World-size-2 tests of the cross-shard path (execution-order resolution + witness merge)."""
import numpy as np
import pytest

from tests import dist_worker


def test_raw_position_arithmetic():
    from ipc_filecoin_proofs_b200.parallel import raw_position_of
    raw = list(range(30))
    D = [3, 4, 10, 29]
    kept = [p for p in raw if p not in D]
    for i, p in enumerate(kept):
        assert raw_position_of(i, D) == p
    assert raw_position_of(0, []) == 0 and raw_position_of(5, [0, 1, 2]) == 8
    from ipc_filecoin_proofs_b200.parallel import raw_positions_of
    assert raw_positions_of(np.arange(len(kept), dtype=np.uint64), np.array(D, dtype=np.uint64)).tolist() == kept
    assert raw_positions_of(np.array([5], dtype=np.uint64), np.array([0, 1, 2], dtype=np.uint64)).tolist() == [8]


def test_cross_shard_protocol_gloo_cpu():
    """2 processes, gloo, no GPU: oracle per rank + host restatement of the device helpers."""
    dist_worker.run(dist_worker.cpu_worker, world=2)


def test_cross_shard_protocol_gloo_cpu_world4():
    """Same at world size 4: shares that start and end inside different message AMTs, three ranks without duplicates."""
    dist_worker.run(dist_worker.cpu_worker, world=4)


@pytest.mark.gpu
def test_cross_shard_engine_two_ranks_one_gpu():
    """2 processes sharing cuda:0 over gloo: the CUDA engine on sharded stores + ipcfp_exec_* helpers."""
    dist_worker.run(dist_worker.gpu_worker, world=2)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_call_over_nccl(world):
    """The in-library protocol (ipcfp_comm_init + ipcfp_generate_event_proof_sharded), one rank per GPU over NCCL, bit-exact against
    the oracle of the whole tipset (proofs incl. message_cid, n_exec, merged witness CID list) and failing together on a fault."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} visible GPUs")
    dist_worker.run(dist_worker.nccl_worker, world=world)
