"""TEST SUPPORT: writes the classic HDF5 on-disk structures (superblock v0, old-style groups = symbol-table message ->
v1 B-tree -> SNOD nodes -> local heap, version-1 object headers, contiguous float32 datasets) that h5py emits by default,
from a nested dict {name: ndarray | dict}.  Written from the published HDF5 file-format specification; NOT produced by
libhdf5 -- see the "parity unpinned" note in tools/h5lite.py."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        at = len(self.b)
        self.b.extend(b"\x00" * n)
        return at

    def put(self, at, data):
        self.b[at:at + len(data)] = data


def _msg(mtype, payload):
    payload = payload + b"\x00" * (-len(payload) % 8)
    return struct.pack("<HHB3x", mtype, len(payload), 0) + payload


def _object_header(buf, messages, continuation_from=None):
    """messages: list of encoded messages.  continuation_from = k puts messages[k:] into a continuation block."""
    if continuation_from is not None and continuation_from < len(messages):
        tail = b"".join(messages[continuation_from:])
        tail_at = buf.alloc(len(tail))
        buf.put(tail_at, tail)
        head = messages[:continuation_from] + [_msg(0x10, struct.pack("<QQ", tail_at, len(tail)))]
        n = len(messages) + 1
    else:
        head, n = messages, len(messages)
    body = b"".join(head)
    at = buf.alloc(16 + len(body))
    buf.put(at, struct.pack("<BBHII4x", 1, 0, n, 1, len(body)) + body)
    return at


def _dataset(buf, arr, variant):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    data_at = buf.alloc(max(arr.nbytes, 1))
    buf.put(data_at, arr.tobytes())
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    # datatype: class 1 (floating point) version 1; bit field: little-endian, IEEE layout; size 4; properties (12 bytes)
    dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    fill = struct.pack("<BBBB", 2, 2, 0, 0)
    layout = struct.pack("<BBQQ", 3, 1, data_at, arr.nbytes)
    mtime = struct.pack("<B3xI", 1, 1506000000)
    msgs = [_msg(0x01, space), _msg(0x03, dtype), _msg(0x05, fill), _msg(0x08, layout), _msg(0x12, mtime)]
    if variant % 3 == 1:
        msgs.insert(2, _msg(0x00, b"\x00" * 8))                   # a NIL message in between
    return _object_header(buf, msgs, continuation_from=3 if variant % 2 else None)


def _group(buf, tree, counter):
    children = {}
    for name, value in tree.items():
        if isinstance(value, dict):
            children[name] = _group(buf, value, counter)
        else:
            counter[0] += 1
            children[name] = _dataset(buf, value, counter[0])
    names = sorted(children)
    heap_data = bytearray(b"\x00" * 8)                            # offset 0: the empty name
    offsets = {}
    for n in names:
        offsets[n] = len(heap_data)
        heap_data += n.encode() + b"\x00"
        heap_data += b"\x00" * (-len(heap_data) % 8)
    seg_at = buf.alloc(len(heap_data))
    buf.put(seg_at, bytes(heap_data))
    heap_at = buf.alloc(32)
    buf.put(heap_at, struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap_data), UNDEF, seg_at))
    snods = []
    for i in range(0, max(len(names), 1), 8):                     # 2K = 8 symbols per node (group leaf K = 4)
        chunk = names[i:i + 8]
        at = buf.alloc(8 + 40 * 8)
        body = struct.pack("<4sBBH", b"SNOD", 1, 0, len(chunk))
        for n in chunk:
            body += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
        buf.put(at, body)
        snods.append((at, offsets[chunk[-1]] if chunk else 0))
    tree_at = buf.alloc(24 + 8 + 16 * len(snods))
    body = struct.pack("<4sBBHQQ", b"TREE", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for at, last_key in snods:
        body += struct.pack("<QQ", at, last_key)
    buf.put(tree_at, body)
    return _object_header(buf, [_msg(0x11, struct.pack("<QQ", tree_at, heap_at))])


def write(path, tree):
    buf = _Buf()
    sb = buf.alloc(96)
    root = _group(buf, tree, [0])
    head = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    head += struct.pack("<QQQQ", 0, UNDEF, len(buf.b), UNDEF)
    head += struct.pack("<QQII16x", 0, root, 0, 0)
    buf.put(sb, head)
    with open(path, "wb") as f:
        f.write(bytes(buf.b))
