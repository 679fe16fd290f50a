#!/usr/bin/env python3
"""Extract golden vectors from the reference tree into tests/golden/ (run in the build container,
where /root/reference exists; the outputs are committed because the GPU box has no /root/reference).

Sources (all real Rust-encoder output):
  crates/examples/examples/issue_stuck.rs:8-9   snapshot (mode 3) + FastUpdates blob GV-1 (mode 4)
  crates/loro/tests/issue_822.bin               snapshot
  crates/loro/tests/issue_import.base64.txt     snapshot
  crates/loro/tests/issue.rs:260                snapshot (inline base64)
Outputs:
  gv1_update.bin                     the 108-byte FastUpdates blob
  snapshot_blocks.json               every change block found in the snapshots' oplog SSTables
                                     {source, key(hex), block(base64)}
  automerge_trace.json.gz            (optional) the editing trace used by text_r.rs benches, re-packed as
                                     [[pos, del, ins], ...] + final text -- only if --trace is given
"""
import base64
import gzip
import json
import os
import re
import struct
import sys

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def lz4_block_decompress(src):
    out = bytearray()
    i = 0
    n = len(src)
    while i < n:
        token = src[i]; i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]; i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]; i += lit
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8); i += 2
        ml = (token & 15) + 4
        if (token & 15) == 15:
            while True:
                b = src[i]; i += 1
                ml += b
                if b != 255:
                    break
        start = len(out) - off
        for k in range(ml):
            out.append(out[start + k])
    return bytes(out)


def lz4_frame_decompress(b):
    assert struct.unpack_from("<I", b, 0)[0] == 0x184D2204
    flg = b[4]
    pos = 6
    if flg & 0x08:
        pos += 8
    if flg & 0x01:
        pos += 4
    pos += 1  # header checksum
    out = bytearray()
    while True:
        (sz,) = struct.unpack_from("<I", b, pos); pos += 4
        if sz == 0:
            break
        raw = sz & 0x80000000
        sz &= 0x7FFFFFFF
        data = b[pos:pos + sz]; pos += sz
        if flg & 0x10:
            pos += 4
        out += data if raw else lz4_block_decompress(data)
    return bytes(out)


def sstable_items(b):
    """docs/encoding.md:175-303 -> list of (key, value)."""
    assert b[:4] == b"LORO", b[:4]
    (meta_off,) = struct.unpack_from("<I", b, len(b) - 4)
    pos = meta_off
    (nblocks,) = struct.unpack_from("<I", b, pos); pos += 4
    metas = []
    for _ in range(nblocks):
        (off,) = struct.unpack_from("<I", b, pos); pos += 4
        (kl,) = struct.unpack_from("<H", b, pos); pos += 2
        first_key = b[pos:pos + kl]; pos += kl
        flags = b[pos]; pos += 1
        is_large = flags >> 7
        comp = flags & 0x7F
        if not is_large:
            (ll,) = struct.unpack_from("<H", b, pos); pos += 2
            pos += ll
        metas.append((off, first_key, is_large, comp))
    items = []
    for i, (off, first_key, is_large, comp) in enumerate(metas):
        end = metas[i + 1][0] if i + 1 < len(metas) else meta_off
        chunk = b[off:end]
        body = chunk[:-4]
        if comp == 1:
            body = lz4_frame_decompress(body)
        if is_large:
            items.append((first_key, body))
            continue
        (cnt,) = struct.unpack_from("<H", body, len(body) - 2)
        offs = struct.unpack_from("<%dH" % cnt, body, len(body) - 2 - 2 * cnt)
        data_end = len(body) - 2 - 2 * cnt
        for k in range(cnt):
            s = offs[k]
            e = offs[k + 1] if k + 1 < cnt else data_end
            ent = body[s:e]
            if k == 0:
                items.append((first_key, ent))
            else:
                cp = ent[0]
                (sl,) = struct.unpack_from("<H", ent, 1)
                key = first_key[:cp] + ent[3:3 + sl]
                items.append((key, ent[3 + sl:]))
    return items


def snapshot_change_blocks(blob):
    assert blob[:4] == b"loro" and blob[20:22] == b"\x00\x03", blob[:22]
    body = blob[22:]
    (olen,) = struct.unpack_from("<I", body, 0)
    oplog = body[4:4 + olen]
    return [(k, v) for k, v in sstable_items(oplog) if len(k) == 12]


def main():
    src = open(f"{REF}/crates/examples/examples/issue_stuck.rs").read()
    snap_b64 = re.search(r'let snapshot = "([^"]+)"', src).group(1)
    upd_b64 = re.search(r'let update = "([^"]+)"', src).group(1)
    open(f"{OUT}/gv1_update.bin", "wb").write(base64.b64decode(upd_b64))
    sources = {
        "issue_stuck.rs:8": base64.b64decode(snap_b64),
        "issue_822.bin": open(f"{REF}/crates/loro/tests/issue_822.bin", "rb").read(),
        "issue_import.base64.txt": base64.b64decode(open(f"{REF}/crates/loro/tests/issue_import.base64.txt").read().strip()),
    }
    issue_rs = open(f"{REF}/crates/loro/tests/issue.rs").read().splitlines()
    for ln in issue_rs[250:275]:
        m = re.search(r'"([A-Za-z0-9+/=]{200,})"', ln)
        if m:
            sources["issue.rs:260"] = base64.b64decode(m.group(1))
            break
    blocks = []
    for name, blob in sources.items():
        for k, v in snapshot_change_blocks(blob):
            blocks.append({"source": name, "key": k.hex(), "block": base64.b64encode(v).decode()})
        print(name, len(blob), "bytes ->", sum(1 for b in blocks if b["source"] == name), "blocks")
    json.dump(blocks, open(f"{OUT}/snapshot_blocks.json", "w"), indent=0)
    if "--trace" in sys.argv:
        tr = json.load(gzip.open(f"{REF}/crates/loro-internal/benches/automerge-paper.json.gz"))
        txns = tr["txns"]
        patches = [p for t in txns for p in t["patches"]]
        out = {"patches": patches, "endContent": tr["endContent"]}
        with gzip.open(f"{OUT}/automerge_trace.json.gz", "wt", compresslevel=9) as f:
            json.dump(out, f, separators=(",", ":"))
        print("trace:", len(patches), "patches")


if __name__ == "__main__":
    main()
