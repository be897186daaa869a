"""oracle/jpeg_decode.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the frame ingest that sits
upstream of the hot path, `convert_image_to_mat` (reference src/face_detection_lite/utils.rs:8-21):

    Mat::from_slice(bytes) -> imgcodecs::imdecode(IMREAD_COLOR) -> cvt_color(COLOR_BGR2RGB)

SURVEY.md section 8f rank 3 ("next" row; nothing in the product path uses this file -- a device decoder is future work, this is
the oracle it will be held to).  The arithmetic lives in OpenCV's bundled libjpeg(-turbo), which is not vendored in the reference
(`opencv` crate 0.93.1 binds the host's libopencv); the published algorithm restated here is the baseline sequential JPEG decode of
ITU-T T.81 with libjpeg's default decompression choices, which is what `imdecode` uses:

  * Huffman entropy decoding, interleaved MCUs, restart intervals (T.81 F.2.2, E.2.4)
  * dequantisation + the accurate integer inverse DCT `jpeg_idct_islow` (jidctint.c: CONST_BITS 13, PASS1_BITS 2, the Loeffler /
    Ligtenberg / Moschytz factorisation), range-limited around +128
  * "fancy" triangle-filter chroma upsampling for 2h2v / 2h1v components (jdsample.c: 3/4-1/4 weights in both directions, the
    alternating +8 / +7 rounding, edge replication at the image -- not the padded block -- borders; plain replication when the
    downsampled component is at most 2 samples wide, as jinit_upsampler chooses)
  * YCbCr -> RGB with the 16-bit fixed-point tables of jdcolor.c

Pinned: bit-exact against `cv2.imdecode` + `cv2.cvtColor(BGR2RGB)` (OpenCV 4.13.0 in this container) on every JPEG under
test_data/ (4:2:0 baseline files, one with a restart interval) and on re-encoded 4:4:4 / 4:2:2 / greyscale variants --
tests/test_oracle_jpeg.py.  Progressive, arithmetic-coded, 12-bit and CMYK files are rejected (ValueError).
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], np.int64)


class _Bits:
    """MSB-first bit reader over entropy-coded data with 0xFF00 byte stuffing removed; stops at markers."""

    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        while self.n <= 24:
            if self.p >= len(self.d):
                b = 0
            else:
                b = self.d[self.p]
                if b == 0xFF:
                    nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nxt == 0:
                        self.p += 2
                    else:
                        b = 0          # a marker: feed zeros, do not advance (T.81 F.2.2.5 / libjpeg's "insufficient data" padding)
                else:
                    self.p += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def peek16(self) -> int:
        if self.n < 16:
            self._fill()
        return (self.acc >> (self.n - 16)) & 0xFFFF

    def skip(self, k: int):
        self.n -= k

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        v = (self.acc >> (self.n - k)) & ((1 << k) - 1)
        self.n -= k
        return v

    def restart(self):
        """Discard the remaining bits, consume the RSTn marker (E.2.4)."""
        self.acc = self.n = 0
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """(code length, symbol) for every 16-bit prefix: a flat lookup (T.81 Annex C code assignment)."""
    lut_len = np.zeros(65536, np.uint8)
    lut_sym = np.zeros(65536, np.uint8)
    code, k = 0, 0
    for length in range(1, 17):
        if code + counts[length - 1] > (1 << length):          # libjpeg: "Bogus Huffman table definition"
            raise ValueError("bad DHT: code lengths over-subscribed")
        for _ in range(counts[length - 1]):
            lo = code << (16 - length)
            hi = lo + (1 << (16 - length))
            lut_len[lo:hi] = length
            lut_sym[lo:hi] = symbols[k]
            k += 1
            code += 1
        code <<= 1
    return lut_len.tolist(), lut_sym.tolist()


def _extend(v: int, t: int) -> int:
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def _idct_islow(coef: np.ndarray) -> np.ndarray:
    """jpeg_idct_islow on dequantised blocks [n,8,8] (row-major: [v][u]) -> samples 0..255 [n,8,8]."""
    F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
             f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)

    def pass_1d(x, shift_in, descale):
        # x: [n,8,8]; transform along axis 1 (index k = frequency), keeping axis 2
        i0, i1, i2, i3, i4, i5, i6, i7 = (x[:, k, :] for k in range(8))
        z1 = (i2 + i6) * F["f0_541"]
        tmp2 = z1 + i6 * (-F["f1_847"])
        tmp3 = z1 + i2 * F["f0_765"]
        tmp0 = (i0 + i4) << shift_in
        tmp1 = (i0 - i4) << shift_in
        tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
        t0, t1, t2, t3 = i7, i5, i3, i1
        z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
        z5 = (z3 + z4) * F["f1_175"]
        t0 = t0 * F["f0_298"]; t1 = t1 * F["f2_053"]; t2 = t2 * F["f3_072"]; t3 = t3 * F["f1_501"]
        z1 = z1 * (-F["f0_899"]); z2 = z2 * (-F["f2_562"]); z3 = z3 * (-F["f1_961"]) + z5; z4 = z4 * (-F["f0_390"]) + z5
        t0 = t0 + z1 + z3; t1 = t1 + z2 + z4; t2 = t2 + z2 + z3; t3 = t3 + z1 + z4
        half = 1 << (descale - 1)
        out = [tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3]
        return np.stack([(o + half) >> descale for o in out], axis=1)

    x = coef.astype(np.int64)
    ws = pass_1d(x, 13, 13 - 2)                                   # columns: frequency index v is axis 1
    rows = pass_1d(ws.transpose(0, 2, 1), 13, 13 + 2 + 3)         # rows: frequency index u; result [n, x, y]
    v = rows.transpose(0, 2, 1) & 0x3FF                           # range_limit[(...) & RANGE_MASK], table centred on +128
    out = np.where(v < 128, v + 128, np.where(v < 512, 255, np.where(v < 896, 0, v - 896)))
    return out.astype(np.uint8)


def _upsample_h2v2(c: np.ndarray) -> np.ndarray:
    """h2v2_fancy_upsample: c [h, w] uint8 (the REAL downsampled size) -> [2h, 2w]."""
    h, w = c.shape
    if w <= 2:                                                     # jinit_upsampler: fancy only when downsampled_width > 2
        return np.repeat(np.repeat(c, 2, axis=0), 2, axis=1)
    ci = c.astype(np.int64)
    above = np.vstack([ci[:1], ci[:-1]])                           # edge rows replicate
    below = np.vstack([ci[1:], ci[-1:]])
    out = np.empty((2 * h, 2 * w), np.int64)
    for v, other in ((0, above), (1, below)):
        s = 3 * ci + other                                         # thiscolsum per column
        last = np.hstack([s[:, :1], s[:, :-1]])
        nxt = np.hstack([s[:, 1:], s[:, -1:]])
        even = (3 * s + last + 8) >> 4
        odd = (3 * s + nxt + 7) >> 4
        if w == 1:
            even[:, 0] = (4 * s[:, 0] + 8) >> 4
            odd[:, 0] = (4 * s[:, 0] + 7) >> 4
        else:
            even[:, 0] = (4 * s[:, 0] + 8) >> 4
            odd[:, -1] = (4 * s[:, -1] + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out.astype(np.uint8)


def _upsample_h2v1(c: np.ndarray) -> np.ndarray:
    """h2v1_fancy_upsample: [h, w] -> [h, 2w]."""
    h, w = c.shape
    if w <= 2:                                                     # jinit_upsampler: fancy only when downsampled_width > 2
        return np.repeat(c, 2, axis=1)
    s = c.astype(np.int64)
    last = np.hstack([s[:, :1], s[:, :-1]])
    nxt = np.hstack([s[:, 1:], s[:, -1:]])
    even = (3 * s + last + 1) >> 2
    odd = (3 * s + nxt + 2) >> 2
    even[:, 0] = s[:, 0]
    odd[:, -1] = s[:, -1]
    out = np.empty((h, 2 * w), np.int64)
    out[:, 0::2] = even
    out[:, 1::2] = odd
    return out.astype(np.uint8)


def _ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    """jdcolor.c ycc_rgb_convert: 16-bit fixed-point tables, arithmetic right shifts, clamp to 0..255."""
    def fix(x):
        return int(x * 65536 + 0.5)
    x = np.arange(256, dtype=np.int64) - 128
    cr_r = (fix(1.40200) * x + 32768) >> 16
    cb_b = (fix(1.77200) * x + 32768) >> 16
    cr_g = -fix(0.71414) * x
    cb_g = -fix(0.34414) * x + 32768
    yi = y.astype(np.int64)
    r = yi + cr_r[cr]
    g = yi + ((cb_g[cb] + cr_g[cr]) >> 16)
    b = yi + cb_b[cb]
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def _exif_orientation(tiff: bytes) -> int:
    """Orientation tag (0x0112) of IFD0, 0 if absent / unreadable."""
    import struct
    if len(tiff) < 8 or tiff[:2] not in (b"II", b"MM"):
        return 0
    e = "<" if tiff[:2] == b"II" else ">"
    off = struct.unpack(e + "I", tiff[4:8])[0]
    if off + 2 > len(tiff):
        return 0
    n = struct.unpack(e + "H", tiff[off:off + 2])[0]
    for k in range(n):
        ent = tiff[off + 2 + 12 * k:off + 14 + 12 * k]
        if len(ent) < 12:
            break
        tag, typ = struct.unpack(e + "HH", ent[:4])
        if tag == 0x0112 and typ == 3:
            return struct.unpack(e + "H", ent[8:10])[0]
    return 0


def entropy_decode(data: bytes) -> dict:
    """Markers + Huffman stage: {"H", "W", "hmax", "vmax", "comps": [{"h", "v", "coef": int64 [blocks_y, blocks_x, 64] in natural
    (row-major) order, quantised, "quant": int64 [64] natural order}, ...]}.  Raises ValueError for anything but 8-bit baseline
    Huffman files with 1 or 3 components and EXIF orientation 1 / none."""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG")
    qt, ht, raw_ht = {}, {}, {}
    frame = None
    restart_interval = 0
    p = 2
    scan = None
    while p < len(data):
        if data[p] != 0xFF:
            raise ValueError("marker expected")
        m = data[p + 1]
        if m == 0xFF:
            p += 1
            continue
        L = (data[p + 2] << 8) | data[p + 3]
        seg = data[p + 4:p + 2 + L]
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                if pq:
                    vals = np.frombuffer(seg[q + 1:q + 129], ">u2").astype(np.int64); q += 129
                else:
                    vals = np.frombuffer(seg[q + 1:q + 65], np.uint8).astype(np.int64); q += 65
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = vals
                qt[tq] = t
        elif m == 0xC0 or m == 0xC1:
            if seg[0] != 8:
                raise ValueError("only 8-bit samples")
            H, W, n = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            if n not in (1, 3):
                raise ValueError("1 or 3 components only")
            frame = dict(H=H, W=W, comps=[dict(id=seg[6 + 3 * k], h=seg[7 + 3 * k] >> 4, v=seg[7 + 3 * k] & 15, tq=seg[8 + 3 * k]) for k in range(n)])
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential Huffman JPEG (SOF0/SOF1)")
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                nsym = sum(counts)
                ht[(tc, th)] = _huff_table(counts, list(seg[q + 17:q + 17 + nsym]))
                raw_ht[(tc, th)] = (bytes(counts), bytes(seg[q + 17:q + 17 + nsym]))
                q += 17 + nsym
        elif m == 0xDD:
            restart_interval = (seg[0] << 8) | seg[1]
        elif m == 0xE1 and seg[:6] == b"Exif\x00\x00" and _exif_orientation(seg[6:]) not in (0, 1):
            # imdecode rotates / mirrors such files after decoding (OpenCV >= 3.1); not restated here
            raise ValueError("EXIF orientation %d: not restated" % _exif_orientation(seg[6:]))
        elif m == 0xDA:
            ns = seg[0]
            scan = [(seg[1 + 2 * k], seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(ns)]
            p += 2 + L
            break
        p += 2 + L
    if frame is None or scan is None:
        raise ValueError("no frame / scan")
    comps = frame["comps"]
    if len(scan) != len(comps):
        raise ValueError("only single-scan (interleaved) files")
    for c, (cid, td, ta) in zip(comps, scan):
        if c["id"] != cid:
            raise ValueError("scan component order")
        c["dc"], c["ac"] = ht[(0, td)], ht[(1, ta)]
        c["raw_dc"], c["raw_ac"] = raw_ht[(0, td)], raw_ht[(1, ta)]
    scan_offset = p
    H, W = frame["H"], frame["W"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-W // (8 * hmax)), -(-H // (8 * vmax))
    for c in comps:
        c["coef"] = np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int64)
        c["pred"] = 0

    # ---- entropy decoding (T.81 F.2.2) ----
    br = _Bits(data, p)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if restart_interval and count and count % restart_interval == 0:
                br.restart()
                for c in comps:
                    c["pred"] = 0
            count += 1
            for c in comps:
                dc_len, dc_sym = c["dc"]
                ac_len, ac_sym = c["ac"]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = c["coef"][my * c["v"] + by, mx * c["h"] + bx]
                        code = br.peek16()
                        ln = dc_len[code]
                        if ln == 0:
                            raise ValueError("bad Huffman code")
                        br.skip(ln)
                        t = dc_sym[code]
                        diff = _extend(br.get(t), t) if t else 0
                        c["pred"] += diff
                        blk[0] = c["pred"]
                        k = 1
                        while k < 64:
                            code = br.peek16()
                            ln = ac_len[code]
                            if ln == 0:
                                raise ValueError("bad Huffman code")
                            br.skip(ln)
                            rs = ac_sym[code]
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16
                                    continue
                                break
                            k += r
                            if k > 63:
                                break
                            blk[ZIGZAG[k]] = _extend(br.get(s), s)
                            k += 1

    return dict(H=H, W=W, hmax=hmax, vmax=vmax, restart_interval=restart_interval, scan_offset=scan_offset,
                comps=[dict(h=c["h"], v=c["v"], coef=c["coef"], quant=qt[c["tq"]], dht_dc=c["raw_dc"], dht_ac=c["raw_ac"]) for c in comps])


def decode_jpeg_rgb(data: bytes) -> np.ndarray:
    """Baseline JPEG bytes -> uint8 [H, W, 3] in RGB order: what `convert_image_to_mat` returns (utils.rs:8-21)."""
    f = entropy_decode(data)
    H, W, hmax, vmax, comps = f["H"], f["W"], f["hmax"], f["vmax"], f["comps"]
    # ---- dequantise, inverse DCT, assemble the component planes (padded to whole blocks) ----
    planes = []
    for c in comps:
        by, bx, _ = c["coef"].shape
        deq = (c["coef"] * c["quant"]).reshape(by * bx, 8, 8)
        px = _idct_islow(deq).reshape(by, bx, 8, 8).transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)
        # the REAL downsampled size (jdmaster.c: ceil(image * samp / max_samp)): upsampling replicates ITS edges, not the padding's
        ch, cw = -(-H * c["v"] // vmax), -(-W * c["h"] // hmax)
        planes.append((px[:ch, :cw], hmax // c["h"], vmax // c["v"]))
    full = []
    for px, eh, ev in planes:
        if eh == 1 and ev == 1:
            up = px
        elif eh == 2 and ev == 2:
            up = _upsample_h2v2(px)
        elif eh == 2 and ev == 1:
            up = _upsample_h2v1(px)
        else:
            raise ValueError("unsupported sampling factors %dx%d" % (eh, ev))
        full.append(up[:H, :W])
    if len(full) == 1:
        return np.repeat(full[0][:, :, None], 3, axis=2)            # IMREAD_COLOR of a greyscale file: grey replicated
    return _ycc_to_rgb(full[0], full[1], full[2])


def convert_image_to_mat(im_bytes: bytes) -> np.ndarray:
    """utils.rs:8-21 (EXIF orientation 1): RGB uint8 [H, W, 3]."""
    return decode_jpeg_rgb(im_bytes)
