"""Index arithmetic of the dense message-AMT walk, restated on the CPU.

`generate_event_proof` plans the walk on the host (csrc/events.cu, `DensePlan`: per round and AMT the first frontier slot, per AMT
the first value slot) and `amt_item_dense` places every child / value by arithmetic on its index instead of a scan. This model
replays exactly those formulas over random AMT shapes (counts, heights incl. non-minimal ones) and shard ranges (as the host
derives them from the receipt range) and checks what the kernel relies on: every slot of every level is written exactly once and
inside the planned size, and the values come out in (AMT, index) order — the reference's execution order before dedup
(events/utils.rs:48-94). The one geometry the plan does not cover (an empty share strictly inside an AMT) is the one the host
hands to the general walk."""
import random


def plan(cnt, hgt, lo, hi):
    namt, rounds = len(cnt), max(hgt) + 1
    fofs = [[0] * namt for _ in range(rounds)]
    ftot = [0] * rounds
    for r in range(rounds):
        run = 0
        for k in range(namt):
            fofs[r][k] = run
            if r > hgt[k]:
                continue
            sh = 3 * (hgt[k] - r + 1)
            run += 1 if r == 0 else ((((hi[k] - 1) >> sh) - (lo[k] >> sh) + 1) if lo[k] < hi[k] else 0)
        ftot[r] = run
    vbase, vb = [], 0
    for k in range(namt):
        vbase.append(vb)
        vb += hi[k] - lo[k]
    return fofs, ftot, vbase, vb


def walk(cnt, hgt, lo, hi):
    fofs, ftot, vbase, nraw = plan(cnt, hgt, lo, hi)
    frontier = {k: (k, hgt[k], 0) for k in range(len(cnt))}      # slot -> (amt, level, base), as k_setup seeds it
    vals = {}
    for r in range(len(ftot)):
        assert sorted(frontier) == list(range(ftot[r]))            # every planned slot written exactly once
        nxt = {}
        for a, level, base in frontier.values():
            sh = 3 * level
            n_exp = min(8, ((cnt[a] - base - 1) >> sh) + 1) if cnt[a] > base else 0
            for j in range(n_exp):
                cb = base + (j << sh)
                if not (cb < hi[a] and cb + (1 << sh) > lo[a]):
                    continue
                if level:
                    d = fofs[r + 1][a] + ((cb >> sh) - (lo[a] >> sh))
                    assert d not in nxt and 0 <= d < ftot[r + 1]
                    nxt[d] = (a, level - 1, cb)
                else:
                    v = vbase[a] + (cb - lo[a])
                    assert v not in vals and 0 <= v < nraw
                    vals[v] = (a, cb)
        frontier = nxt
    assert sorted(vals) == list(range(nraw))
    seq = [vals[i] for i in range(nraw)]
    assert seq == sorted(seq)
    return nraw


def shard_ranges(cnt, n_receipts, rank, world):
    """Per-AMT index range of a shard, as the host computes it (events.cu, `h_rng`), clipped to [0, count]."""
    total = sum(cnt)
    rl, rh = n_receipts * rank // world, n_receipts * (rank + 1) // world
    glo, ghi = total * rl // n_receipts, total * rh // n_receipts
    lo, hi, a0 = [], [], 0
    for c in cnt:
        a1 = a0 + c
        l = glo - a0 if glo > a0 else 0
        h = ghi - a0 if ghi > a0 else 0
        if world == 1:
            l, h = 0, 1 << 64
        elif glo >= a1 and not (a1 == a0 and glo == a0):
            l = h = 0
        elif ghi >= a1:
            h = 1 << 64
        h = max(h, l)
        l = min(l, c)
        h = max(l, min(h, c))
        lo.append(l)
        hi.append(h)
        a0 = a1
    return lo, hi


def test_dense_walk_positions():
    rng = random.Random(20260922)
    walked = skipped = 0
    for _ in range(2500):
        namt = rng.randint(1, 6)
        cnt = [rng.choice([0, 1, 7, 8, 9, 63, 64, 65, 100, 511, 512, 513, rng.randint(0, 5000)]) for _ in range(namt)]
        hgt = []
        for c in cnt:
            h = 0
            while 8 ** (h + 1) < max(c, 1):
                h += 1
            hgt.append(h + rng.choice([0, 0, 0, 1, 2]))               # roots may be taller than needed
        world = rng.choice([1, 2, 3, 8])
        n_receipts = rng.randint(1, 10000)
        owned = 0
        for rank in range(world):
            lo, hi = shard_ranges(cnt, n_receipts, rank, world)
            owned += sum(h - l for l, h in zip(lo, hi))
            if any(l == h and l > 0 for l, h in zip(lo, hi)):        # the host leaves this geometry to the general walk
                skipped += 1
                continue
            assert walk(cnt, hgt, lo, hi) == sum(h - l for l, h in zip(lo, hi))
            walked += 1
        assert owned == sum(cnt)                                     # the shards tile the raw list
    assert walked > 5000 and skipped < walked
