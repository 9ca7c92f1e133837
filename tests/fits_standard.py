"""An independent reader of FITS primary images, written from the FITS standard (v4.0, sections 3.1, 4.1-4.4, 5.1-5.3) and the
ESO HIERARCH keyword convention — not from the repo's own writer or its test reader (slicer_b200/host.py:read_fits), with which
it shares no code.  It VALIDATES while it reads: anything a conforming reader (CFITSIO, astropy.io.fits — the consumers in
Lens/kslicer.py:39-40,84-86) could reject raises AssertionError.  Keyword lookup is case-insensitive and HIERARCH-transparent,
as in astropy (which is how kslicer.py finds `DLLOW` although the writer calls the key `DlLOW`)."""
import re
import struct

BLOCK = 2880
CARD = 80
_KEY_RE = re.compile(r"^[A-Z0-9_-]{0,8}$")
_INT_RE = re.compile(r"^[+-]?\d+$")
_REAL_RE = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([ED][+-]?\d+)?$")


class Header:
    def __init__(self):
        self.cards = []  # (keyword as written, value, comment)

    def __getitem__(self, key):
        k = key.upper()
        hits = [v for (kw, v, _) in self.cards if kw.upper() == k]
        if not hits:
            raise KeyError(key)
        return hits[0]

    def __contains__(self, key):
        return any(kw.upper() == key.upper() for (kw, _, _) in self.cards)

    def keywords(self):
        return [kw for (kw, _, _) in self.cards]


def _parse_value(field, where):
    """Value field of a card (after the value indicator) -> (python value, comment)."""
    f = field
    stripped = f.lstrip()
    if stripped.startswith("'"):
        # character string: starts with a quote, '' is an embedded quote, closing quote required (4.2.1)
        i = f.index("'") + 1
        out = []
        while True:
            assert i < len(f), f"{where}: unterminated string"
            if f[i] == "'":
                if i + 1 < len(f) and f[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(f[i])
            i += 1
        rest = f[i + 1:]
        value = "".join(out).rstrip()
    else:
        head, sep, tail = f.partition("/")
        rest = sep + tail
        tok = head.strip()
        assert tok != "", f"{where}: empty value"
        if tok in ("T", "F"):
            value = tok == "T"
        elif _INT_RE.match(tok):
            value = int(tok)
        else:
            assert _REAL_RE.match(tok), f"{where}: '{tok}' is neither logical, integer nor real (4.2.3, 4.2.4)"
            value = float(tok.replace("D", "E"))
    rest = rest.strip()
    comment = ""
    if rest:
        assert rest.startswith("/"), f"{where}: text after the value must be a comment introduced by '/' (4.1.2.3)"
        comment = rest[1:].strip()
    return value, comment


def read_primary_image(path):
    raw = open(path, "rb").read()
    assert len(raw) % BLOCK == 0, "a FITS file is a whole number of 2880-byte blocks (3.1)"
    hdr = Header()
    off = 0
    end = False
    while not end:
        assert off + BLOCK <= len(raw), "header without END"
        block = raw[off:off + BLOCK]
        off += BLOCK
        for c in range(BLOCK // CARD):
            card_b = block[c * CARD:(c + 1) * CARD]
            assert all(32 <= b <= 126 for b in card_b), "header cards hold printable ASCII only (4.1.1)"
            card = card_b.decode("ascii")
            where = f"card {len(hdr.cards)} '{card.rstrip()}'"
            if end:
                assert card.strip() == "", "the rest of the last header block is blank (4.3.1, END)"
                continue
            if card.startswith("END") and card[3:].strip() == "":
                assert card[:8] == "END     ", "END occupies columns 1-8 (4.4.1)"
                end = True
                continue
            if card.startswith("HIERARCH "):
                # ESO convention: free-format keyword tokens up to '=', then an ordinary value field
                body = card[9:]
                assert "=" in body, f"{where}: HIERARCH card without '='"
                kw, _, val = body.partition("=")
                kw = kw.strip()
                assert kw and all(33 <= ord(ch) <= 126 for ch in kw.replace(" ", "")), where
                value, comment = _parse_value(val, where)
                hdr.cards.append((kw, value, comment))
                continue
            kw = card[:8].rstrip()
            assert " " not in kw, f"{where}: embedded blank in keyword"
            if card[8:10] == "= ":
                assert _KEY_RE.match(kw), f"{where}: keyword must be upper-case letters, digits, '-' or '_' in columns 1-8 (4.1.2.1)"
                value, comment = _parse_value(card[10:], where)
                if isinstance(value, (bool, int, float)) and "'" not in card[10:].split("/")[0]:
                    # fixed format is mandatory for the mandatory keywords: right-justified in columns 11-30 (4.2.3)
                    if kw in ("SIMPLE", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2"):
                        assert card[10:30].strip() == card[10:30].lstrip() and card[29] != " ", f"{where}: mandatory keyword not in fixed format"
                hdr.cards.append((kw, value, comment))
            else:
                assert kw in ("COMMENT", "HISTORY", ""), f"{where}: no value indicator and not a commentary keyword"
    kws = hdr.keywords()
    # mandatory keywords of a primary header, in order (4.4.1.1)
    assert kws[:3] == ["SIMPLE", "BITPIX", "NAXIS"], kws[:5]
    assert hdr["SIMPLE"] is True
    naxis = hdr["NAXIS"]
    assert kws[3:3 + naxis] == [f"NAXIS{i + 1}" for i in range(naxis)]
    bitpix = hdr["BITPIX"]
    assert bitpix in (8, 16, 32, 64, -32, -64)
    assert len(set(k.upper() for k in kws)) == len(kws), "a keyword appears twice"
    shape = [hdr[f"NAXIS{i + 1}"] for i in range(naxis)]
    n = 1
    for s in shape:
        n *= s
    nbytes = n * abs(bitpix) // 8
    padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
    assert len(raw) == off + padded, "data array followed by nothing but its padding (3.3.2, 5.1)"
    assert not any(raw[off + nbytes:]), "the data padding is zero filled (3.3.2)"
    assert bitpix == -32, "this reader handles BITPIX -32 images"
    data = struct.unpack(f">{n}f", raw[off:off + nbytes])  # big-endian IEEE-754 (5.3)
    # NAXIS1 is the fastest axis (5.1): rows of NAXIS1 pixels
    rows = [list(data[r * shape[0]:(r + 1) * shape[0]]) for r in range(shape[1])] if naxis == 2 else list(data)
    return hdr, rows
