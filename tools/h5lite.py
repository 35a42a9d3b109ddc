"""Minimal reader for the HDF5 files Keras 2.x writes with ``model.save_weights`` (the reference's
``model_weight.h5`` / ``model_best_weight.h5``, agent/model.py:82-101) -- pure Python + numpy, because neither h5py nor
libhdf5 exists in the target image.

NOT ON THE PRODUCT PATH (round 2): the self-play worker takes float32 blobs only (tools/export_keras_weights.py writes them
on the trainer side, where Keras/h5py exist) and refuses a model directory that holds nothing but h5 files.  This reader is
kept as a stand-alone conversion tool (`python tools/h5lite.py <model_weight.h5> <out.rzblob.npy> [filters blocks value_fc]`).

PARITY UNPINNED: no HDF5 file written by libhdf5 is available where this was built, so the reader has only been checked
against files produced by this repo's own writer of the same on-disk structures (tests/support/h5_write.py).  It reads
the classic layout h5py emits by default (``libver='earliest'``): superblock version 0 or 1, old-style groups (symbol
table message -> v1 B-tree -> SNOD nodes -> local heap), version-1 object headers with continuation blocks, contiguous
or compact dataset layouts, IEEE little-endian float32 / float64 data.  Anything else (new-style groups, chunked or
filtered datasets, other datatypes) raises ``H5FormatError`` -- nothing is guessed; use tools/export_keras_weights.py on
the trainer side in that case.

    datasets = read_datasets("data/model/model_best_weight.h5")    # {"conv2d_1/conv2d_1/kernel:0": ndarray, ...}
    layers = keras_layers_from_h5(datasets)                        # [(layer_name, class_name, [arrays])] for weights_from_keras_layers
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class _File:
    def __init__(self, data):
        self.d = data
        if data[:8] != SIGNATURE:
            raise H5FormatError("not an HDF5 file (signature at offset 0 expected)")
        ver = data[8]
        if ver not in (0, 1):
            raise H5FormatError(f"superblock version {ver} not supported (only the classic versions 0 and 1)")
        if data[13] != 8 or data[14] != 8:
            raise H5FormatError("only 8-byte offsets and lengths are supported")
        at = 24 if ver == 0 else 28          # v1 has indexed-storage K (2) + reserved (2) after the group K values
        self.base, _free, _eof, _drv = struct.unpack_from("<QQQQ", data, at)
        ste = at + 32                         # root group symbol table entry
        _name_off, self.root_header, cache_type = struct.unpack_from("<QQI", data, ste)
        self.root_scratch = struct.unpack_from("<QQ", data, ste + 24) if cache_type == 1 else None

    def u(self, fmt, at):
        if at + struct.calcsize(fmt) > len(self.d):
            raise H5FormatError("address beyond the end of the file")
        return struct.unpack_from(fmt, self.d, at)

    # ---- version-1 object header: list of (type, flags, payload offset, payload size) ---------------------
    def messages(self, addr):
        addr += self.base
        version, _r, nmsg, _refs, hsize = self.u("<BBHII", addr)
        if version != 1:
            raise H5FormatError(f"object header version {version} not supported (only version 1)")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            at, size = blocks.pop(0)
            end = at + size
            while at + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u("<HHB", at)
                body = at + 8
                if mtype == 0x10:            # continuation: offset, length of another block of messages
                    off, length = self.u("<QQ", body)
                    blocks.append((off + self.base, length))
                out.append((mtype, flags, body, msize))
                at = body + msize
        return out

    # ---- old-style group: children as {name: object header address} ----------------------------------------
    def group_children(self, addr):
        st = [m for m in self.messages(addr) if m[0] == 0x11]
        if not st:
            raise H5FormatError("group without a symbol table message (new-style groups are not supported)")
        btree, heap = self.u("<QQ", st[0][2])
        return self._children(btree, heap)

    def _children(self, btree, heap):
        sig, _v, _r, _seg_size, _free, seg = self.u("<4sB3sQQQ", heap + self.base)
        if sig != b"HEAP":
            raise H5FormatError("local heap signature expected")
        names = {}
        self._walk(btree, seg + self.base, names, depth=0)
        return names

    def _walk(self, node, heap_data, names, depth):
        if depth > 32:
            raise H5FormatError("B-tree too deep")
        node += self.base
        sig = self.d[node:node + 4]
        if sig == b"TREE":
            ntype, _level, used = self.u("<BBH", node + 4)
            if ntype != 0:
                raise H5FormatError("group B-tree node expected")
            at = node + 24               # signature 4, type 1, level 1, entries 2, left sibling 8, right sibling 8
            for i in range(used):        # key0 child0 key1 child1 ... keyN
                child = self.u("<Q", at + 8 + i * 16)[0]
                self._walk(child, heap_data, names, depth + 1)
        elif sig == b"SNOD":
            _v, _r, count = self.u("<BBH", node + 4)
            for i in range(count):
                e = node + 8 + i * 40
                name_off, header = self.u("<QQ", e)
                end = self.d.index(b"\x00", heap_data + name_off)
                names[self.d[heap_data + name_off:end].decode("utf-8")] = header
        else:
            raise H5FormatError("B-tree or symbol-table node expected")

    def is_group(self, addr):
        return any(m[0] == 0x11 for m in self.messages(addr))

    # ---- dataset ---------------------------------------------------------------------------------------------
    def read_dataset(self, addr):
        shape = dtype = layout = None
        for mtype, _flags, at, size in self.messages(addr):
            if mtype == 0x01:            # dataspace
                ver, rank, _fl = self.u("<BBB", at)
                dims_at = at + (8 if ver == 1 else 4)
                if ver not in (1, 2):
                    raise H5FormatError(f"dataspace version {ver} not supported")
                shape = tuple(self.u("<Q", dims_at + 8 * i)[0] for i in range(rank))
            elif mtype == 0x03:          # datatype
                cls_ver, bits0, _b1, _b2, tsize = self.u("<BBBBI", at)
                if cls_ver & 0x0F != 1 or bits0 & 1 or tsize not in (4, 8):
                    raise H5FormatError("only little-endian IEEE float32 / float64 datasets are supported")
                dtype = np.dtype("<f4" if tsize == 4 else "<f8")
            elif mtype == 0x08:          # data layout
                ver, cls = self.u("<BB", at)
                if ver != 3:
                    raise H5FormatError(f"data layout version {ver} not supported (only version 3)")
                if cls == 1:
                    layout = ("contiguous",) + self.u("<QQ", at + 2)
                elif cls == 0:
                    layout = ("compact", at + 4, self.u("<H", at + 2)[0])
                else:
                    raise H5FormatError("chunked datasets are not supported (Keras writes contiguous weights)")
            elif mtype == 0x0B:
                raise H5FormatError("filtered (compressed) datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise H5FormatError("dataset without dataspace / datatype / layout message")
        n = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            off, size = layout[1], layout[2]
            if off == UNDEF:
                return np.zeros(shape, dtype)          # storage never allocated: all fill values
            off += self.base
        else:
            off, size = layout[1], layout[2]
        if size < n * dtype.itemsize or off + n * dtype.itemsize > len(self.d):
            raise H5FormatError("dataset storage smaller than its extent")
        return np.frombuffer(self.d, dtype, n, off).reshape(shape).astype(np.float32)


def read_datasets(path):
    """All datasets of the file as {"group/sub/name": float32 ndarray}."""
    with open(path, "rb") as f:
        h = _File(f.read())
    out = {}

    def visit(addr, prefix, depth):
        if depth > 16:
            raise H5FormatError("group nesting too deep")
        for name, child in sorted(h.group_children(addr).items()):
            full = f"{prefix}{name}"
            if h.is_group(child):
                visit(child, full + "/", depth + 1)
            else:
                out[full] = h.read_dataset(child)

    visit(h.root_header, "", 0)
    return out


_KERAS_WEIGHT_ORDER = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}


def keras_layers_from_h5(datasets):
    """Keras ``save_weights`` stores layer ``L``'s weight ``W`` at ``L/L/W:0`` (the weight name ``L/W:0`` below the layer's
    group ``L``).  Returns ``[(layer_name, class_name, [arrays in Keras' weight order])]`` for
    ``agent.model.weights_from_keras_layers``; the class is inferred from the weight names."""
    layers = {}
    for path, arr in datasets.items():
        parts = path.split("/")
        if len(parts) < 2:
            continue
        layer, wname = parts[0], parts[-1].split(":")[0]
        if wname not in _KERAS_WEIGHT_ORDER:
            raise H5FormatError(f"unexpected weight name {path!r}")
        layers.setdefault(layer, {})[wname] = arr
    out = []
    for layer, ws in layers.items():
        if "gamma" in ws:
            cls = "BatchNormalization"
        elif ws["kernel"].ndim == 4:
            cls = "Conv2D"
        else:
            cls = "Dense"
        out.append((layer, cls, [ws[k] for k in sorted(ws, key=_KERAS_WEIGHT_ORDER.get)]))
    return out


def blob_from_keras_h5(mc, path):
    """model_weight.h5 of the reference (Keras save_weights) -> the float32 blob rz_net_load_weights takes."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reversi-alpha-zero_b200"))
    from reversi_zero_b200.agent import model as M
    return M.weights_to_blob(mc, M.weights_from_keras_layers(mc, keras_layers_from_h5(read_datasets(path))))


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reversi-alpha-zero_b200"))
    from reversi_zero_b200.agent import model as M
    f, b, v = (int(x) for x in sys.argv[3:6]) if len(sys.argv) >= 6 else (256, 10, 256)
    np.save(sys.argv[2], blob_from_keras_h5(M.ModelConfig(cnn_filter_num=f, res_layer_num=b, value_fc_size=v), sys.argv[1]))
