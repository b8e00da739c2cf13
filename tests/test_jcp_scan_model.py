"""Host model of the chain stage of k_jcp_rows (csrc/segment.cu): along an image row the JCP recurrence
is a composition of finite maps on the nine states (outcome of entry q-2, outcome of entry q-1). The
model scans the maps in chunks of 32 with a carried state, and must agree with the plain left-to-right
walk over the state plane for random rows - including queued neighbours that a vote masks out and gaps
between entries. The kernel holds a map as nine bytes and composes with PRMT (byte permute, selector
nibble 8 = "replicate the sign"): `prmt` below restates the PTX instruction, and the byte-form
composition must equal the plain table composition on every map the scan produces."""
import numpy as np

IDENTITY = 0x876543210


def compose(later: int, earlier: int) -> int:
    r = 0
    for x in range(9):
        idx = (earlier >> (4 * x)) & 15
        r |= ((later >> (4 * idx)) & 15) << (4 * x)
    return r


def entry_map(entry: int, prev_is_w2: bool) -> int:
    m = 0
    for x in range(9):
        a, b = divmod(x, 3)
        out = 0
        if entry & 0x200:
            s11 = b if entry & 0x800 else 0
            s10 = (b if prev_is_w2 else a) if entry & 0x400 else 0
            out = 2 if (entry >> (s10 * 3 + s11)) & 1 else 1
        m |= (b * 3 + out) << (4 * x)
    return m


def walk(ws, entries):
    """The reference order: one pixel after the other, neighbours read from the plane."""
    plane = {}
    for w, e in zip(ws, entries):
        out = 0
        if e & 0x200:
            s10 = plane.get(w - 2, 0) if e & 0x400 else 0
            s11 = plane.get(w - 1, 0) if e & 0x800 else 0
            out = 2 if (e >> (s10 * 3 + s11)) & 1 else 1
        plane[w] = out
    return [plane[w] for w in ws]


def scan(ws, entries, chunk=32):
    outs, carry = [], 0
    for c0 in range(0, len(ws), chunk):
        maps = []
        for q in range(c0, min(c0 + chunk, len(ws))):
            prev_is_w2 = q > 0 and ws[q - 1] + 2 == ws[q]
            maps.append(entry_map(entries[q], prev_is_w2))
        # inclusive scan (Hillis-Steele, as the warp does with shuffles)
        off = 1
        while off < chunk:
            maps = [compose(maps[i], maps[i - off]) if i >= off else maps[i] for i in range(len(maps))]
            off <<= 1
        states = [(m >> (4 * carry)) & 15 for m in maps]
        outs += [s % 3 for s in states]
        carry = states[-1]
    return outs


def random_row(rng, width=300, density=0.8):
    ws = np.flatnonzero(rng.random(width) < density)
    queued = set(int(w) for w in ws)
    entries = []
    for w in ws:
        e = int(rng.integers(0, 512))
        if rng.random() < 0.9:
            e |= 0x200                       # decidable
        # a dependency needs a queued neighbour, but a queued neighbour may be masked out of the vote
        if (w - 2) in queued and rng.random() < 0.8:
            e |= 0x400
        if (w - 1) in queued and rng.random() < 0.8:
            e |= 0x800
        entries.append(e)
    return [int(w) for w in ws], entries


def test_identity_and_composition_order():
    rng = np.random.default_rng(1)
    f = entry_map(0x200 | 0x800 | int(rng.integers(0, 512)), False)
    g = entry_map(0x200 | 0x400 | 0x800 | int(rng.integers(0, 512)), False)
    assert compose(f, IDENTITY) == f and compose(IDENTITY, f) == f
    for x in range(9):
        assert (compose(g, f) >> (4 * x)) & 15 == (g >> (4 * ((f >> (4 * x)) & 15))) & 15


def test_scan_equals_walk_on_random_rows():
    rng = np.random.default_rng(2)
    for trial in range(60):
        ws, entries = random_row(rng, width=int(rng.integers(1, 400)), density=float(rng.uniform(0.2, 1.0)))
        if ws:
            assert scan(ws, entries) == walk(ws, entries), trial


# ---- the kernel's representation: nine bytes (r0, r1, r2), composition by byte permutes (segment.cu: jcp_compose)
def prmt(a: int, b: int, sel: int) -> int:
    """PTX prmt.b32, default mode: nibble i of sel picks byte (n & 7) of the 8 bytes {a, b}; bit 3 of the
    nibble replicates that byte's sign bit over the result byte instead."""
    src = [(a >> (8 * i)) & 0xff for i in range(4)] + [(b >> (8 * i)) & 0xff for i in range(4)]
    r = 0
    for i in range(4):
        n = (sel >> (4 * i)) & 0xf
        v = src[n & 7]
        if n & 8:
            v = 0xff if v & 0x80 else 0x00
        r |= v << (8 * i)
    return r


def to_bytes(m: int):
    img = [(m >> (4 * x)) & 15 for x in range(9)]
    return (img[0] | img[1] << 8 | img[2] << 16 | img[3] << 24, img[4] | img[5] << 8 | img[6] << 16 | img[7] << 24, img[8])


def pack(r0, r1, r2):
    t0, t1 = r0 | (r0 >> 4), r1 | (r1 >> 4)
    return prmt(t0, t1, 0x6420), r2


def compose_bytes(later, e0, e1):
    l8 = prmt(later[2], 0, 0x0000)
    k = 0x80808080
    r0 = prmt(later[0], later[1], e0 & 0xffffffff) | (prmt(k, k, e0) & l8)
    r1 = prmt(later[0], later[1], e0 >> 16) | (prmt(k, k, e0 >> 16) & l8)
    r2 = (prmt(later[0], later[1], e1) | (prmt(k, k, e1) & l8)) & 0xff
    return (r0, r1, r2)


def apply_bytes(m, x):
    return (m[2] if x == 8 else prmt(m[0], m[1], x)) & 0xff


def test_byte_permute_composition_equals_table_composition():
    rng = np.random.default_rng(3)
    assert to_bytes(IDENTITY) == (0x03020100, 0x07060504, 8)
    for trial in range(400):
        # arbitrary maps on the nine states (the scan's partial products are not entry maps any more)
        f = sum(int(rng.integers(0, 9)) << (4 * x) for x in range(9))
        g = sum(int(rng.integers(0, 9)) << (4 * x) for x in range(9))
        n0, n1 = pack(*to_bytes(f))
        assert n0 == f & 0xffffffff and n1 == f >> 32
        got = compose_bytes(to_bytes(g), n0, n1)
        assert got == to_bytes(compose(g, f)), trial
        for x in range(9):
            assert apply_bytes(got, x) == (compose(g, f) >> (4 * x)) & 15
