"""Seeded synthetic tumor/normal long-read windows in the reference's npz row format
``[sequenceList, ReadIDs, flank_5, flank_3, TDRecord]`` (the rows written by
src/SomTDDetector_AimDatFetch.py:120 and read back by src/SVscope.py:210-212).

Definitions follow SURVEY.md §8(d): what reaches ``poa`` is the read sub-sequence between the
50-bp flank anchors, so "read length" here is the length of those strings.

    C1  one window: 10 kb reference body + 2x50 flanks, 30 normal + 30 tumor reads, 15 tumor
        reads carry a 2 kb deletion, 5 % ONT-like noise (40 % sub / 30 % ins / 30 % del).
    C2  INS/DEL windows: type p=1/2, SV length log-uniform 50-2000, body 5-15 kb uniform,
        30 tumor + 30 normal, VAF uniform 0.2-0.6, 5 % noise; window i uses seed 1000+i.
    C3  tandem-repeat INS: 60+60 reads x 20 kb, motif 2-60 bp, tumor subset with extra copies,
        10 % noise.
    C4  C2 distribution, seeds 0.., with a 1 % C3-like heavy tail.
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
FLANK = 50


def _rand_seq(rng, n):
    return rng.integers(0, 4, size=int(n), dtype=np.uint8)


def _to_str(codes) -> str:
    return BASES[codes].tobytes().decode()


def noisy_copy(rng, template: np.ndarray, err: float, mix=(0.4, 0.3, 0.3)) -> np.ndarray:
    """i.i.d. per-base substitution / insertion-after / deletion."""
    n = template.shape[0]
    if n == 0 or err <= 0:
        return template.copy()
    u = rng.random(n)
    p_sub, p_ins, p_del = (err * m for m in mix)
    is_sub = u < p_sub
    is_ins = (u >= p_sub) & (u < p_sub + p_ins)
    is_del = (u >= p_sub + p_ins) & (u < p_sub + p_ins + p_del)
    base = template.copy()
    base[is_sub] = (base[is_sub] + rng.integers(1, 4, size=int(is_sub.sum()), dtype=np.uint8)) % 4
    counts = np.ones(n, dtype=np.int64)
    counts[is_del] = 0
    counts[is_ins] = 2
    out = np.repeat(base, counts)
    # the second copy of an "insertion" position becomes a random base
    ends = np.cumsum(counts)
    ins_pos = ends[is_ins] - 1
    out[ins_pos] = rng.integers(0, 4, size=ins_pos.shape[0], dtype=np.uint8)
    return out


def _window_record(chrom, start, ref_codes, reads_codes, tags, sample="S"):
    ref = _to_str(ref_codes)
    seqs = [ref] + [_to_str(r) for r in reads_codes]
    ids = np.array([f"{sample}_{t}|r{i}" for i, t in enumerate(tags)])
    rec = f"{chrom}\t{start}\t{start + len(ref) - 2 * FLANK}"
    return [seqs, ids, ref[:FLANK], ref[-FLANK:], rec]


def make_sv_window(seed: int, body_len: int, sv_type: str, sv_len: int, n_tumor=30, n_normal=30,
                   n_carriers=15, err=0.05, chrom="chr1", start=None, sv_offset=None):
    """Tumor reads first, then normal reads (ReadIDs order = alignment order)."""
    rng = np.random.default_rng(seed)
    ref = _rand_seq(rng, body_len + 2 * FLANK)
    if sv_type == "DEL":
        sv_len = int(min(sv_len, max(1, body_len - 200)))
        lo, hi = FLANK + 50, FLANK + body_len - sv_len - 50
        off = int(sv_offset) if sv_offset is not None else int(rng.integers(lo, max(lo + 1, hi)))
        alt = np.concatenate([ref[:off], ref[off + sv_len:]])
    elif sv_type == "INS":
        lo, hi = FLANK + 50, FLANK + body_len - 50
        off = int(sv_offset) if sv_offset is not None else int(rng.integers(lo, max(lo + 1, hi)))
        alt = np.concatenate([ref[:off], _rand_seq(rng, sv_len), ref[off:]])
    else:
        raise ValueError(sv_type)
    reads, tags = [], []
    for i in range(n_tumor):
        reads.append(noisy_copy(rng, alt if i < n_carriers else ref, err))
        tags.append("tumor")
    for _ in range(n_normal):
        reads.append(noisy_copy(rng, ref, err))
        tags.append("normal")
    if start is None:
        start = 1_000_000 + (seed % 100_000) * 20_000
    return _window_record(chrom, start, ref, reads, tags)


def make_c1(seed: int = 1, body_len: int = 10_000, sv_len: int = 2_000, n_tumor=30, n_normal=30,
            n_carriers=15, err=0.05):
    return make_sv_window(seed, body_len, "DEL", sv_len, n_tumor, n_normal, n_carriers, err,
                          sv_offset=FLANK + (body_len - sv_len) // 2)


def make_c2_window(index: int, body_range=(5_000, 15_000), sv_range=(50, 2_000), depth=30, err=0.05,
                   seed_base: int = 1000):
    seed = seed_base + index
    rng = np.random.default_rng([seed, 7])
    sv_type = "INS" if rng.random() < 0.5 else "DEL"
    sv_len = int(round(np.exp(rng.uniform(np.log(sv_range[0]), np.log(sv_range[1])))))
    body = int(rng.integers(body_range[0], body_range[1] + 1))
    vaf = rng.uniform(0.2, 0.6)
    carriers = int(max(3, round(vaf * depth)))
    return make_sv_window(seed, body, sv_type, sv_len, depth, depth, carriers, err,
                          chrom=f"chr{1 + index % 22}", start=1_000_000 + index * 20_000)


def make_c2(n_windows: int = 1000, **kw):
    return [make_c2_window(i, **kw) for i in range(n_windows)]


def make_c3(seed: int = 3, total_len: int = 20_000, n_tumor=60, n_normal=60, n_carriers=30,
            err=0.10, extra_copies=None):
    rng = np.random.default_rng(seed)
    motif_len = int(rng.integers(2, 61))
    motif = _rand_seq(rng, motif_len)
    copies = max(1, (total_len - 2 * FLANK) // motif_len)
    body = np.tile(motif, copies)
    f5, f3 = _rand_seq(rng, FLANK), _rand_seq(rng, FLANK)
    ref = np.concatenate([f5, body, f3])
    k = int(extra_copies) if extra_copies is not None else int(rng.integers(5, 40))
    alt = np.concatenate([f5, np.tile(motif, copies + k), f3])
    reads, tags = [], []
    for i in range(n_tumor):
        reads.append(noisy_copy(rng, alt if i < n_carriers else ref, err))
        tags.append("tumor")
    for _ in range(n_normal):
        reads.append(noisy_copy(rng, ref, err))
        tags.append("normal")
    return _window_record("chr3", 3_000_000 + seed * 50_000, ref, reads, tags)


def make_c4(n_windows: int, heavy_tail: float = 0.01, first: int = 0, **kw):
    """C2 distribution with seeds first..first+n-1 plus a ``heavy_tail`` fraction of C3-like windows."""
    out = []
    for i in range(first, first + n_windows):
        pick = np.random.default_rng([i, 11]).random()
        if pick < heavy_tail:
            out.append(make_c3(seed=100_000 + i))
        else:
            out.append(make_c2_window(i, seed_base=0, **kw))
    return out


def make_small_window(seed: int, body_len=300, sv_len=60, n_tumor=8, n_normal=8, n_carriers=5,
                      err=0.05, sv_type="DEL"):
    """Scaled-down C1 used by parity tests (oracle finishes in milliseconds)."""
    return make_sv_window(seed, body_len, sv_type, sv_len, n_tumor, n_normal, n_carriers, err)


def save_npz(path: str, windows) -> None:
    """Same container the reference dumps: ``np.savez(path, DatSet=array_of_object_rows)``
    (SomTDDetector_AimDatFetch.py:173,183)."""
    arr = np.empty(len(windows), dtype=object)
    for i, w in enumerate(windows):
        arr[i] = np.array(w, dtype=object)
    np.savez(path, DatSet=arr)


def load_npz(path: str):
    dat = np.load(path, allow_pickle=True)["DatSet"]
    return [list(dat[i]) for i in range(dat.shape[0])]


def window_cost(window) -> float:
    """Cost model used for sharding: POA cells + Myers cells estimate."""
    seqs = window[0]
    n = max(1, len(seqs) - 1)
    lbar = float(np.mean([len(s) for s in seqs])) if seqs else 0.0
    eps = 0.035
    return (n + 1) * lbar * (lbar + n * lbar * eps) * 0.5 + n * n * lbar * lbar / 64.0
