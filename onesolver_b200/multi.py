"""Multi-GPU plumbing: trajectory sharding and the single end-of-run best-energy gather.

Trajectories are independent (reference annealing.hpp:85-86), so ranks shard the global
trajectory ids [first_try, first_try + tries) with Q replicated and no data-path collective;
the only exchange is one all-gather of {energy, global id, packed state} per rank
(~0.5 KB at N=4096), after which every rank picks min energy then min id -- the
std::min_element rule of annealing.hpp:134 applied across shards.
"""
import numpy as np


def shard(num_tries, world, rank):
    """Contiguous id range of `rank`: (first_try, count). Remainder goes to the low ranks.
    count is 0 for the high ranks when num_tries < world: such a rank skips its local anneal and
    contributes encode_best(inf, 2**64 - 1, zeros) to the gather (osa_multi_anneal does the same)."""
    base, rem = divmod(num_tries, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def encode_best(energy, index, state):
    bits = np.packbits(np.asarray(state, dtype=np.uint8), bitorder="little")
    buf = np.zeros(16 + bits.size, dtype=np.uint8)
    buf[:8] = np.frombuffer(np.float64(energy).tobytes(), dtype=np.uint8)
    buf[8:16] = np.frombuffer(np.uint64(index).tobytes(), dtype=np.uint8)
    buf[16:] = bits
    return buf


def decode_best(rows, n):
    """rows: [world][16 + ceil(n/8)] uint8 -> (energy, index, state) of the global winner."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    energies = rows[:, :8].copy().view(np.float64).ravel()
    ids = rows[:, 8:16].copy().view(np.uint64).ravel()
    live = [r for r in range(rows.shape[0]) if ids[r] != np.uint64(0xFFFFFFFFFFFFFFFF)]
    if not live:
        raise ValueError("no rank produced a result")
    k = min(live, key=lambda r: (energies[r], ids[r]))
    state = np.unpackbits(rows[k, 16:], bitorder="little")[:n]
    return float(energies[k]), int(ids[k]), state


def gather_best(dist, torch, energy, index, state, device):
    """One collective (NCCL on GPUs, gloo in the CPU tests) gathering every rank's local best."""
    n = len(state)
    buf = torch.from_numpy(encode_best(energy, index, state)).to(device)
    world = dist.get_world_size()
    out = torch.empty(world * buf.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, buf)
    rows = out.cpu().numpy().reshape(world, -1)
    return decode_best(rows, n)
