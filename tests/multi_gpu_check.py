"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): the sharded hot path against the
single-GPU path computed on every rank from the same inputs (qio_b200.selfcheck), several Jacobi cycles with the
mailbox epoch wrapping around, and the stress test of the tagged-word signalling over NVLink."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    from qio_b200 import selfcheck
    from qio_b200.peers import PeerGroup

    peers = PeerGroup()
    sizes = tuple(int(x) for x in os.environ.get('QIO_CHECK_N', '13,24').split(','))
    res = selfcheck.multi_gpu_parity(peers, sizes=sizes, verbose=True)
    assert res['ok'], res['failures']
    # the epoch wrap (mailboxes cleared between two barriers) in the middle of a run of cycles
    peers.EPOCHS = peers.struct.epoch + 3
    res2 = selfcheck.multi_gpu_parity(peers, sizes=sizes[:1], extra_cycles=6)
    assert res2['ok'], res2['failures']
    assert peers.wraps >= 1, 'the epoch did not wrap'
    peers.EPOCHS = 4095
    stress = selfcheck.mailbox_stress(peers, iterations=int(os.environ.get('QIO_STRESS_ITERS', '4000')))
    assert stress['ok'] and stress['words_checked'] == 4000 * 256 * world or os.environ.get('QIO_STRESS_ITERS'), stress
    if rank == 0:
        print(f"multi-GPU check ok: sizes={sizes}, world={world}, {res['checks']} + {res2['checks']} checks; mailbox stress: "
              f"{stress['stores_per_rank']} tagged stores per rank, {stress['words_checked']} words verified, "
              f"{stress['mismatches']} mismatches, {stress['timeouts']} time-outs", flush=True)
    peers.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
