"""Adversarial inputs for the per-read device logic (host build, lane-group width 1) against the
oracle: count profiles that no simulator produces -- plateaus at and around the H, D and repeat
levels, error-like dips of K-1 positions, counts up to 32767, +-1 noise -- on sequences full of
homopolymer / di- / tri-nucleotide stretches.  The input domain is the reference's defined one:
counts >= 1 (a k-mer of a read occurs at least once in the read set; the reference crashes on zeros)
and no low-complexity run of 127 units or more (beyond its 127 cap the reference reads cells it
never wrote; the device flags such reads CPG_ST_LONG_RUN = 32 and they are skipped here)."""
import numpy as np
import pytest


def rand_seq(rng, n):
    parts, tot = [], 0
    while tot < n:
        kind = rng.integers(0, 5)
        if kind <= 1:
            p = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(3, 60))).tolist())
        else:
            u = bytes(rng.choice(list(b"ACGT"), size=int(kind - 1)).tolist())
            p = (u * 100)[:int(rng.integers(len(u), 90))] + bytes(rng.choice(list(b"ACGT"), size=3).tolist())
        parts.append(p)
        tot += len(p)
    return b"".join(parts)[:n]


def rand_counts(rng, n, H, D, R):
    c = np.zeros(n, np.int64)
    i = 0
    level = int(rng.choice([H, D, D, H, R + 5, 1, 3 * D]))
    while i < n:
        L = int(rng.integers(1, 300))
        mode = rng.integers(0, 10)
        if mode < 5:
            level = int(rng.choice([H, D, D, H, max(1, H + int(rng.integers(-3, 4))), max(1, D + int(rng.integers(-5, 6)))]))
        elif mode == 5:
            level = int(rng.integers(1, 4))                    # error-like
        elif mode == 6:
            level = int(rng.integers(R - 3, R + 40))           # around the repeat threshold
        elif mode == 7:
            level = int(rng.integers(200, 32767))              # high-copy repeat
        elif mode == 8:
            level = int(rng.integers(1, 3))
        else:
            level = max(1, level + int(rng.integers(-6, 7)))
        seg = np.full(min(L, n - i), level, np.int64)
        if rng.random() < 0.5:
            seg = seg + rng.integers(-1, 2, size=len(seg))
        if rng.random() < 0.2 and len(seg) > 45:               # a dip of K-1 positions (an error in the read)
            a = int(rng.integers(0, len(seg) - 40))
            seg[a:a + 39] = int(rng.integers(1, 3))
        c[i:i + len(seg)] = seg
        i += len(seg)
    return np.clip(c, 1, 32767).astype(np.uint16)


@pytest.mark.parametrize("seed", [11, 12])
def test_adversarial_profiles_match_the_oracle(kit, hostsim, seed):
    rng = np.random.default_rng(seed)
    sim = kit.simulate(seed=5, genome_len=30000, cov=20., het=0.01, len_mean=3000)       # for a valid histogram / model
    total = flagged = 0
    for cov_opt, read_len in ((0, 20000), (30, 20000), (12, 8000), (100, 25000)):
        om = kit.oracle_model(sim, cov_opt, read_len)
        gm = kit.gpu_model_from_sim(hostsim, sim, cov_opt, read_len)
        H, D, R = om.cov[2], om.cov[3], om.cov[1]
        ow = kit.OracleWork(clean=True)
        for r in range(120):
            n = int(rng.integers(1, 3000))
            s = rand_seq(rng, n + 39)
            c = rand_counts(rng, n, H, D, R)
            a = ow.classify(om, s, c)
            st, b = kit.hostsim_classify(gm, s, c, 2)
            total += 1
            if st & 32:
                flagged += 1
                continue
            assert (st & ~128) == 0 and a == b, (seed, cov_opt, r, n, st)
    assert flagged < total // 10


@pytest.mark.parametrize("group,small_caps", [(4, 0), (8, 0), (16, 5), (0, 4)])
def test_adversarial_profiles_on_lane_groups_and_the_retry_path(kit, hostsim, group, small_caps):
    """The same inputs on the 32-thread warp emulation with lane groups of 4 / 8 / 16 (group = 0: the
    width-1 build), and with interval tables of a few entries in the first attempt, so that most
    reads take the abort paths of the phase code and are classified again with full-size tables."""
    lib = kit.hostsim32_lib() if group else hostsim
    rng = np.random.default_rng(100 + group + small_caps)
    sim = kit.simulate(seed=5, genome_len=30000, cov=20., het=0.01, len_mean=3000)
    if group:
        assert lib.hs_set_group(group) == 0
    lib.hs_set_small_caps(small_caps)
    try:
        for cov_opt, read_len in ((0, 20000), (12, 8000)):
            om = kit.oracle_model(sim, cov_opt, read_len)
            gm = kit.gpu_model_from_sim(lib, sim, cov_opt, read_len)
            H, D, R = om.cov[2], om.cov[3], om.cov[1]
            ow = kit.OracleWork(clean=True)
            for r in range(80):
                n = int(rng.integers(1, 1500))
                s = rand_seq(rng, n + 39)
                c = rand_counts(rng, n, H, D, R)
                a = ow.classify(om, s, c)
                st, b = kit.hostsim_classify(gm, s, c, 2, lib=lib)
                if st & 32:
                    continue
                assert (st & ~128) == 0 and a == b, (group, small_caps, cov_opt, r, n, st)
        if small_caps:
            assert lib.hs_retries() > 40
    finally:
        lib.hs_set_small_caps(0)
        if group:
            lib.hs_set_group(32)
