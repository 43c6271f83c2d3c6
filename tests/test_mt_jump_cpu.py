"""CPU check of the mt19937 jump-ahead table the device generator uses (csrc/mt_jump_table.h, written by
tools/gen_mt_jump.py): the committed polynomials must advance a generator state by exactly
MT_JUMP_CHUNK_WORDS * 2^m words, as sequential generation does, and the chunk plan of cdlrm_rngdev_raw
(restated here) must reproduce numpy's own MT19937 stream (= torch's CPU generator, main_no_ddp.py:183-185)."""
import importlib.util
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gen_mt_jump", os.path.join(ROOT, "tools", "gen_mt_jump.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)


def _table():
    txt = open(os.path.join(ROOT, "cdlrm_b200", "csrc", "mt_jump_table.h")).read()
    cb = int(re.search(r"#define MT_JUMP_BLOCKS_PER_CHUNK (\d+)", txt).group(1))
    levels = int(re.search(r"#define MT_JUMP_LEVELS (\d+)", txt).group(1))
    rows = re.findall(r"\{((?:0x[0-9a-f]{8}u,?)+)\}", txt)
    assert len(rows) == levels
    polys = []
    for r in rows:
        words = [int(w[:-1], 16) for w in r.split(",") if w]
        assert len(words) == 624
        polys.append(sum(w << (32 * i) for i, w in enumerate(words)))
    return cb, polys


def _temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000))
    return y ^ (y >> np.uint32(18))


def test_committed_polynomials_jump_like_sequential_generation():
    cb, polys = _table()
    assert cb == G.BLOCKS_PER_CHUNK and len(polys) == G.LEVELS
    b0 = G.mt_blocks(G.seed_state(99), 5)[-1]
    window = np.concatenate([b0, G.mt_blocks(b0, 33).reshape(-1)])
    cur = b0
    for m in range(2):                                   # 4096 and 8192 blocks (higher levels: gen_mt_jump.py --verify-levels)
        want = G.mt_blocks(b0, cb << m)[-1]
        got = G.jump_words(window, polys[m])
        assert np.array_equal(got[1:], want[1:]) and (int(got[0]) >> 31) == (int(want[0]) >> 31)
    # the higher levels are squares of the lower ones modulo phi: g_{m+1} = g_m^2 mod phi
    # (phi itself is re-derived by the generator script; here only the degree bound is checked)
    assert all(p.bit_length() <= 19937 for p in polys)


def test_chunk_plan_reproduces_the_sequential_stream():
    """The host logic of cdlrm_rngdev_raw in numpy: position in the current block from the draw count, pending
    words, chunk start states by jumps of the level-0/1 polynomials, blocks per chunk -- against numpy's MT19937
    (init_genrand seeding, the generator torch.manual_seed uses)."""
    cb, polys = _table()
    seed = 4242
    ref = np.random.MT19937()
    ref._legacy_seeding(seed)
    state = G.seed_state(seed)
    pos, drawn = 624, 0
    x = state.copy()
    for n_words in (1000, 2 * cb * 624 + 12345):         # a short request, then one of three chunks from mid-block
        want = ref.random_raw(n_words).astype(np.uint32)
        r0 = 624 - pos
        nb = (n_words - r0 + 623) // 624 if n_words > r0 else 0
        n_chunks = (nb + cb - 1) // cb
        out = np.empty(n_words, dtype=np.uint32)
        head = min(r0, n_words)
        out[:head] = _temper(x[pos:pos + head])
        if nb:
            starts = [x.copy()]
            for j in range(1, n_chunks):                 # chunk j from chunk j - 2^m, m = highest set bit of j
                m = j.bit_length() - 1
                src = starts[j - (1 << m)]
                window = np.concatenate([src, G.mt_blocks(src, 33).reshape(-1)])
                starts.append(G.jump_words(window, polys[m]))
            last = None
            for j in range(n_chunks):
                q0, q1 = j * cb + 1, min(nb, (j + 1) * cb)
                blocks = G.mt_blocks(starts[j], q1 - q0 + 1)
                base = r0 + (q0 - 1) * 624
                flat = _temper(blocks.reshape(-1))
                cnt = min(flat.size, n_words - base)
                out[base:base + cnt] = flat[:cnt]
                last = blocks[-1]
            x, pos = last, n_words - (r0 + (nb - 1) * 624)
        else:
            pos += n_words
        assert np.array_equal(out, want)
        drawn += n_words
        assert pos == ((drawn - 1) % 624) + 1
