"""Minimal pure-Python HDF5 reader / writer with the slice of the h5py API the reference uses.

Why: the reference reads its volumes, weights and predictions through h5py (utils/ImageDataset.py:47-85,
Network/PatchHandler3D.py:122-142, utils/prediction_utils.py:15-28, Network/h5util.py:5-23, Keras
``load_weights`` / ``save`` at predictor.py:61 and TrainerController.py:356,394); h5py / libhdf5 are not
part of this image.  When h5py *is* importable it is used instead (``open_file``).

Reader: superblock v0/v1, version-1 object headers (+ continuations), symbol-table groups (v1 B-tree + local
heap), contiguous / compact / chunked (v1 chunk B-tree) layouts, deflate + shuffle filters, fixed-point, IEEE
float and fixed-length string datatypes, simple attributes.  This covers files written by h5py with its
default ``libver='earliest'`` (the reference's data files and Keras weight files).
Writer: new files only (superblock v0, one symbol-table node per group, contiguous little-endian datasets);
"append" and "resize" are implemented by loading the file and rewriting it, which is what the reference's
small result / quicksave files need.
"""
import os
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


# ================================================================================================
# reader
# ================================================================================================
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise OSError("not an HDF5 file (signature)")
        ver = buf[8]
        if ver not in (0, 1):
            raise OSError(f"HDF5 superblock version {ver} not supported by the shim (needs libver='earliest' files)")
        if buf[13] != 8 or buf[14] != 8:
            raise OSError("only 8-byte offsets/lengths supported")
        pos = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", buf, pos)[0]
        # root symbol table entry follows base, free-space, eof, driver addresses
        ent = pos + 32
        _, self.root_header, cache_type = struct.unpack_from("<QQI", buf, ent)
        self.root_scratch = struct.unpack_from("<QQ", buf, ent + 24) if cache_type == 1 else None

    # ---- object headers ------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, bytes)] of a version-1 object header (continuations followed)."""
        b = self.b
        if b[addr:addr + 4] == b"OHDR":
            raise OSError("version-2 object headers are not supported by the shim")
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise OSError(f"object header version {ver} not supported")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:                                   # continuation
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                out.append((mtype, flags, data))
        return out

    # ---- groups ----------------------------------------------------------------------------------
    def group_entries(self, header_addr):
        """{name: object header address} of a symbol-table group."""
        btree = heap = None
        for mtype, _, data in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", data, 0)
        if btree is None:
            return None
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise OSError("bad local heap")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        out = {}

        def name_at(off):
            s = heap_data + off
            e = b.index(b"\x00", s)
            return b[s:e].decode()

        def walk(node):
            if b[node:node + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, node + 6)[0]
                for i in range(n):
                    noff, haddr = struct.unpack_from("<QQ", b, node + 8 + i * 40)
                    out[name_at(noff)] = haddr
                return
            if b[node:node + 4] != b"TREE":
                raise OSError("bad group B-tree node")
            _, level, used = struct.unpack_from("<BBH", b, node + 4)
            pos = node + 24
            for i in range(used):
                child = struct.unpack_from("<Q", b, pos + 8)[0]      # key_i (8), child_i (8)
                walk(child)
                pos += 16
        if btree != UNDEF:
            walk(btree)
        return out

    # ---- datasets -------------------------------------------------------------------------------
    @staticmethod
    def _dtype(data):
        cls = data[0] & 0x0F
        bits0 = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise OSError(f"HDF5 datatype class {cls} not supported by the shim")

    @staticmethod
    def _shape(data):
        ver, rank, flags = data[0], data[1], data[2]
        pos = 8 if ver == 1 else 4
        dims = struct.unpack_from(f"<{rank}Q", data, pos) if rank else ()
        maxdims = None
        if flags & 1:
            maxdims = struct.unpack_from(f"<{rank}Q", data, pos + 8 * rank)
        return tuple(int(d) for d in dims), maxdims

    def dataset_info(self, header_addr):
        info = {"filters": [], "attrs": {}}
        for mtype, _, data in self.messages(header_addr):
            if mtype == 0x01:
                info["shape"], info["maxshape"] = self._shape(data)
            elif mtype == 0x03:
                info["dtype"] = self._dtype(data)
            elif mtype == 0x08:
                info["layout"] = data
            elif mtype == 0x0B:
                info["filters"] = self._filters(data)
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(data)
                    info["attrs"][k] = v
                except OSError:
                    pass
        return info

    @staticmethod
    def _filters(data):
        ver, n = data[0], data[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, nlen, _, ncv = struct.unpack_from("<HHHH", data, pos)
            pos += 8
            if ver == 1 or fid >= 256:
                pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cvals = struct.unpack_from(f"<{ncv}I", data, pos)
            pos += 4 * ncv
            if ver == 1 and ncv % 2:
                pos += 4
            out.append((fid, cvals))
        return out

    def _attribute(self, data):
        ver = data[0]
        if ver != 1:
            raise OSError("attribute version")
        nsz, dsz, ssz = struct.unpack_from("<HHH", data, 2)
        pos = 8
        name = data[pos:pos + nsz].split(b"\x00")[0].decode()
        pos += (nsz + 7) // 8 * 8
        dt = self._dtype(data[pos:pos + dsz])
        pos += (dsz + 7) // 8 * 8
        shape, _ = self._shape(data[pos:pos + ssz])
        pos += (ssz + 7) // 8 * 8
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(data, dtype=dt, count=n, offset=pos).reshape(shape)
        return name, arr.copy()

    def read_dataset(self, info):
        shape, dt = info["shape"], info["dtype"]
        lay = info["layout"]
        ver, cls = lay[0], lay[1]
        if ver != 3:
            raise OSError(f"data layout version {ver} not supported by the shim")
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:                                                  # compact
            size = struct.unpack_from("<H", lay, 2)[0]
            return np.frombuffer(lay[4:4 + size], dtype=dt, count=n).reshape(shape).copy()
        if cls == 1:                                                  # contiguous
            addr, size = struct.unpack_from("<QQ", lay, 2)
            if addr == UNDEF:
                return np.zeros(shape, dt)
            return np.frombuffer(self.b, dtype=dt, count=n, offset=addr).reshape(shape).copy()
        if cls != 2:
            raise OSError("unknown layout class")
        rank1 = lay[2]                                                # dataset rank + 1
        btree = struct.unpack_from("<Q", lay, 3)[0]
        cdims = struct.unpack_from(f"<{rank1}I", lay, 11)
        chunk = tuple(cdims[:-1])
        out = np.zeros(shape, dt)
        if btree == UNDEF or n == 0:
            return out
        b = self.b

        def walk(node):
            if b[node:node + 4] != b"TREE":
                raise OSError("bad chunk B-tree node")
            _, level, used = struct.unpack_from("<BBH", b, node + 4)
            pos = node + 24
            ksz = 8 + 8 * rank1
            for i in range(used):
                csize, fmask = struct.unpack_from("<II", b, pos)
                offs = struct.unpack_from(f"<{rank1}Q", b, pos + 8)
                child = struct.unpack_from("<Q", b, pos + ksz)[0]
                if level > 0:
                    walk(child)
                else:
                    raw = bytes(b[child:child + csize])
                    for k, (fid, cvals) in reversed(list(enumerate(info["filters"]))):
                        if fmask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cvals[0] if cvals else dt.itemsize
                            a = np.frombuffer(raw, np.uint8)
                            m = a.size // es
                            raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                        elif fid == 3:                                # fletcher32: strip the checksum
                            raw = raw[:-4]
                        else:
                            raise OSError(f"HDF5 filter {fid} not supported by the shim")
                    blk = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunk))).reshape(chunk)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:-1], chunk, shape))
                    out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                pos += ksz + 8
        walk(btree)
        return out


# ================================================================================================
# in-memory tree + writer
# ================================================================================================
class Dataset:
    """numpy-backed dataset node with the h5py.Dataset members the reference touches."""

    def __init__(self, file, name, data, maxshape=None, attrs=None):
        self._file, self.name = file, name
        self._data = np.asarray(data)
        self.maxshape = maxshape
        self.attrs = dict(attrs or {})

    shape = property(lambda self: self._data.shape)
    dtype = property(lambda self: self._data.dtype)
    ndim = property(lambda self: self._data.ndim)

    def __len__(self):
        return self._data.shape[0]

    def __getitem__(self, idx):
        return self._data[idx]

    def __setitem__(self, idx, value):
        self._file._require_write()
        self._data[idx] = value
        self._file._dirty = True

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def resize(self, size, axis=None):
        self._file._require_write()
        new = list(self._data.shape)
        if axis is None:
            new = list(size)
        else:
            new[axis] = int(size)
        out = np.zeros(new, self._data.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(new, self._data.shape))
        out[sl] = self._data[sl]
        self._data = out
        self._file._dirty = True


class Group:
    def __init__(self, file, name):
        self._file, self.name = file, name
        self._items = {}
        self.attrs = {}

    def _resolve(self, path, create=False):
        node = self
        parts = [p for p in path.split("/") if p]
        for p in parts[:-1]:
            if p not in node._items:
                if not create:
                    raise KeyError(path)
                node._items[p] = Group(self._file, node.name.rstrip("/") + "/" + p)
            node = node._items[p]
        return node, (parts[-1] if parts else "")

    def __contains__(self, path):
        try:
            node, leaf = self._resolve(path)
        except KeyError:
            return False
        return leaf in node._items

    def __getitem__(self, path):
        node, leaf = self._resolve(path)
        if leaf not in node._items:
            raise KeyError(f"Unable to open object (object '{path}' doesn't exist)")
        return node._items[leaf]

    def get(self, path, default=None):
        return self[path] if path in self else default

    def keys(self):
        return self._items.keys()

    def items(self):
        return self._items.items()

    def values(self):
        return self._items.values()

    def __iter__(self):
        return iter(self._items)

    def __len__(self):
        return len(self._items)

    def create_group(self, path):
        self._file._require_write()
        node, leaf = self._resolve(path, create=True)
        g = Group(self._file, node.name.rstrip("/") + "/" + leaf)
        node._items[leaf] = g
        self._file._dirty = True
        return g

    def require_group(self, path):
        return self[path] if path in self else self.create_group(path)

    def create_dataset(self, path, data=None, shape=None, dtype=None, maxshape=None, compression=None, **_):
        """`compression` is accepted and ignored: the shim writes contiguous data (the reference writes gzip,
        utils/prediction_utils.py:15-28; h5py reads either).  Said once per process so nobody expects small files."""
        global _WARNED_COMPRESSION
        if compression is not None and not _WARNED_COMPRESSION:
            _WARNED_COMPRESSION = True
            print(f"h5io: compression={compression!r} requested; the built-in HDF5 writer stores datasets uncompressed "
                  "(install h5py to get the reference's gzip files)")
        self._file._require_write()
        if data is None:
            data = np.zeros(shape, dtype or np.float32)
        data = np.array(data, dtype=dtype) if dtype is not None else np.array(data)
        node, leaf = self._resolve(path, create=True)
        if leaf in node._items:
            raise ValueError(f"Unable to create dataset (name already exists): {path}")
        d = Dataset(self._file, node.name.rstrip("/") + "/" + leaf, data, maxshape)
        node._items[leaf] = d
        self._file._dirty = True
        return d

    def visititems(self, fn, _prefix=""):
        for k, v in self._items.items():
            p = f"{_prefix}{k}"
            r = fn(p, v)
            if r is not None:
                return r
            if isinstance(v, Group):
                r = v.visititems(fn, p + "/")
                if r is not None:
                    return r
        return None


_WARNED_COMPRESSION = False


class File(Group):
    """h5py.File work-alike: modes 'r', 'w', 'a' (and 'r+'); context manager; flushes on close."""

    def __init__(self, path, mode="r"):
        super().__init__(self, "/")
        self.filename, self.mode = path, mode
        self._dirty = False
        exists = os.path.exists(path)
        if mode == "r" and not exists:
            raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{path}')")
        if mode in ("r", "r+", "a") and exists:
            with open(path, "rb") as f:
                self._load(f.read())
        elif mode == "r+":
            raise FileNotFoundError(path)
        if mode == "w":
            self._dirty = True

    def _require_write(self):
        if self.mode == "r":
            raise OSError("file is open read-only")

    def _load(self, buf):
        rd = _Reader(memoryview(buf).toreadonly() if not isinstance(buf, bytes) else buf)

        def fill(group, header_addr):
            entries = rd.group_entries(header_addr)
            for name, addr in (entries or {}).items():
                sub = rd.group_entries(addr)
                if sub is not None:
                    g = Group(self, group.name.rstrip("/") + "/" + name)
                    group._items[name] = g
                    for mtype, _, data in rd.messages(addr):
                        if mtype == 0x0C:
                            try:
                                k, v = rd._attribute(data)
                                g.attrs[k] = v
                            except OSError:
                                pass
                    fill(g, addr)
                else:
                    info = rd.dataset_info(addr)
                    if "dtype" not in info or "shape" not in info:
                        continue
                    group._items[name] = Dataset(self, group.name.rstrip("/") + "/" + name, rd.read_dataset(info),
                                                 info.get("maxshape"), info["attrs"])
        for mtype, _, data in rd.messages(rd.root_header):      # attributes of the root group (keras_version, backend, ...)
            if mtype == 0x0C:
                try:
                    k, v = rd._attribute(data)
                    self.attrs[k] = v
                except OSError:
                    pass
        fill(self, rd.root_header)

    def flush(self):
        if self._dirty and self.mode != "r":
            tmp = self.filename + ".tmp"
            with open(tmp, "wb") as f:
                f.write(_serialize(self))
            os.replace(tmp, self.filename)
            self._dirty = False

    def close(self):
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 0x1F, 0x00)
        elif size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 0x3F, 0x00)
        elif size == 2:
            props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
            bits = (0x20, 0x0F, 0x00)
        else:
            raise TypeError(dt)
        return struct.pack("<BBBBI", 0x11, bits[0], bits[1], bits[2], size) + props
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)
    if dt.kind == "b":
        return struct.pack("<BBBBI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8)
    raise TypeError(f"dtype {dt} not supported by the HDF5 shim writer")


def _space_msg(shape, maxshape=None):
    rank = len(shape)
    flags = 1 if maxshape is not None else 0
    out = struct.pack("<BBBBI", 1, rank, flags, 0, 0) + struct.pack(f"<{rank}Q", *shape)
    if maxshape is not None:
        out += struct.pack(f"<{rank}Q", *[UNDEF if m is None else int(m) for m in maxshape])
    return out


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _attr_msg(name, value):
    arr = np.asarray(value)
    if arr.dtype.kind == "U":
        arr = arr.astype("S")
    if arr.dtype.kind == "O":
        arr = np.asarray([str(x).encode() for x in arr.ravel()]).reshape(arr.shape)
    nm = name.encode() + b"\x00"
    dt = _dtype_msg(arr.dtype)
    sp = _space_msg(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + arr.tobytes()
    return _msg(0x0C, body)


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4 + body


def _serialize(root):
    """Lay the tree out depth-first: [superblock][root header]...; every group = header + heap + B-tree + SNOD."""
    chunks = []
    pos = [0]

    def alloc(data):
        data = _pad8(data)
        addr = pos[0]
        chunks.append(data)
        pos[0] += len(data)
        return addr

    def patch(addr, data):
        # chunks are appended in address order; find and overwrite
        off = 0
        for i, c in enumerate(chunks):
            if off == addr:
                assert len(c) >= len(data)
                chunks[i] = data + c[len(data):]
                return
            off += len(c)
        raise AssertionError("patch address")

    LEAF_K = 512      # symbol-table node capacity 2K entries: one SNOD per group

    def write_dataset(d):
        arr = np.ascontiguousarray(d._data)
        if arr.dtype.kind == "U":
            arr = arr.astype("S")
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        daddr = alloc(raw) if raw else UNDEF
        msgs = [_msg(0x01, _space_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype), flags=1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),
                _msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]
        for k, v in d.attrs.items():
            msgs.append(_attr_msg(k, v))
        return alloc(_object_header(msgs))

    def write_group(g):
        children = []
        for name in sorted(g._items, key=lambda s: s.encode()):
            node = g._items[name]
            addr = write_group(node)[0] if isinstance(node, Group) else write_dataset(node)
            children.append((name, addr))
        if len(children) > 2 * LEAF_K:
            raise OSError("too many entries in one group for the HDF5 shim writer")
        # local heap: offset 0 holds the empty string, names follow (8-byte aligned)
        heap = bytearray(b"\x00" * 8)
        noffs = []
        for name, _ in children:
            noffs.append(len(heap))
            heap += _pad8(name.encode() + b"\x00")
        heap_data_size = max(len(heap), 16)
        heap = bytes(heap) + b"\x00" * (heap_data_size - len(heap))
        heap_data_addr = alloc(heap)
        heap_addr = alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, heap_data_size, UNDEF, heap_data_addr))
        # free list head UNDEF encoded as lengths-size all-ones is not valid: use "no free block" = 1 past end
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(children)))
        for (name, addr), noff in zip(children, noffs):
            snod += struct.pack("<QQII", noff, addr, 0, 0) + b"\x00" * 16
        snod += b"\x00" * (8 + 2 * LEAF_K * 40 - len(snod))
        snod_addr = alloc(bytes(snod))
        last_key = noffs[-1] if noffs else 0
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if children else 0, UNDEF, UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, last_key)
        tree += b"\x00" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(tree))
        tree_addr = alloc(tree)
        msgs = [_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in g.attrs.items():
            msgs.append(_attr_msg(k, v))
        return alloc(_object_header(msgs)), tree_addr, heap_addr

    sb_size = 24 + 32 + 40
    alloc(b"\x00" * sb_size)
    root_addr, tree_addr, heap_addr = write_group(root)
    eof = pos[0]
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
    patch(0, sb)
    return b"".join(chunks)


# ================================================================================================
# front door
# ================================================================================================
_READ_CACHE = {}      # path -> (mtime_ns, size, File): parsed read-only files (the training iterator reopens its
_READ_CACHE_MAX = 8   # two data files for every sample, PatchHandler3D.py:122-142)


class _SharedReadOnly:
    """Context-manager view of a cached read-only File (closing it does not drop the cache entry)."""

    def __init__(self, f):
        self._f = f

    def __getattr__(self, k):
        return getattr(self._f, k)

    def __getitem__(self, k):
        return self._f[k]

    def __contains__(self, k):
        return k in self._f

    def __iter__(self):
        return iter(self._f)

    def __len__(self):
        return len(self._f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


def open_file(path, mode="r"):
    """h5py.File when h5py is installed, the shim otherwise (read-only opens share one parsed copy per file)."""
    try:
        import h5py
        if not getattr(h5py, "__shim__", False):
            return h5py.File(path, mode)
    except ImportError:
        pass
    if mode != "r":
        _READ_CACHE.pop(os.path.abspath(path), None)
        return File(path, mode)
    key = os.path.abspath(path)
    st = os.stat(key)
    hit = _READ_CACHE.get(key)
    if hit is None or hit[0] != st.st_mtime_ns or hit[1] != st.st_size:
        if len(_READ_CACHE) >= _READ_CACHE_MAX:
            _READ_CACHE.pop(next(iter(_READ_CACHE)))
        hit = (st.st_mtime_ns, st.st_size, File(key, "r"))
        _READ_CACHE[key] = hit
    return _SharedReadOnly(hit[2])


def install_as_h5py():
    """Register the shim as ``h5py`` in sys.modules when the real package is absent, so reference modules that
    ``import h5py`` (ImageDataset.py, PatchHandler3D.py, h5util.py, prediction_utils.py) run unmodified."""
    import sys
    import types
    try:
        import h5py  # noqa: F401
        return False
    except ImportError:
        m = types.ModuleType("h5py")
        m.File, m.Group, m.Dataset = File, Group, Dataset
        m.__shim__ = True
        sys.modules["h5py"] = m
        return True


# ---- Keras weight files --------------------------------------------------------------------------
def load_keras_weights(path, variable_names, report=None):
    """{name: array} for names like 'conv3d_7/kernel' from a Keras HDF5 file: either a weights file
    (<layer>/<layer>/kernel:0) or a full model file (model_weights/<layer>/<layer>/kernel:0), as written by
    `model.save(path)` / `save_weights` (TrainerController.py:80,356) and read at predictor.py:61.

    Matching.  (1) By name -- Keras auto-names the Conv3D layers conv3d, conv3d_1, ... in creation order, which is this
    package's table order.  (2) When the names do not line up (the process that wrote the file had built other Conv3D
    layers before, so its names start at e.g. conv3d_37 -- Keras' own `load_weights` matches by topology, not by name,
    and still loads such a file): the file's conv layers are sorted by their numeric suffix (= creation order) and
    mapped one to one onto the table, every shape checked.  `report` (a dict, optional) receives the scheme used."""
    found = {}
    with open_file(path, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f

        def visit(p, node):
            if hasattr(node, "shape") and not hasattr(node, "keys"):
                parts = p.split("/")
                leaf = parts[-1].split(":")[0]
                if leaf in ("kernel", "bias") and len(parts) >= 2:
                    found[f"{parts[-2]}/{leaf}"] = np.asarray(node[...] if node.shape else node[()])
            return None
        root.visititems(visit)
    missing = [n for n in variable_names if n not in found]
    if not missing:
        if report is not None:
            report["scheme"] = "by name"
        return {n: found[n] for n in variable_names}

    # fallback: creation order by numeric suffix
    def suffix(layer):
        tail = layer.rsplit("_", 1)[-1]
        return int(tail) if tail.isdigit() else 0
    want_layers = []
    for n in variable_names:
        ly = n.split("/")[0]
        if ly not in want_layers:
            want_layers.append(ly)
    have_layers = sorted({k.split("/")[0] for k in found if k.split("/")[0].split("_")[0].startswith("conv3d")}, key=suffix)
    if len(have_layers) != len(want_layers):
        raise KeyError(f"{path}: missing weights {missing[:4]}{'...' if len(missing) > 4 else ''} and the file holds "
                       f"{len(have_layers)} conv layers where {len(want_layers)} are needed")
    ren = dict(zip(want_layers, have_layers))
    out = {}
    for n in variable_names:
        ly, leaf = n.split("/")
        key = f"{ren[ly]}/{leaf}"
        if key not in found:
            raise KeyError(f"{path}: layer {ren[ly]} (for {ly}) has no {leaf}")
        out[n] = found[key]
    if report is not None:
        report["scheme"] = f"by creation order (file layers {have_layers[0]}..{have_layers[-1]} -> {want_layers[0]}..{want_layers[-1]})"
    print(f"load_keras_weights: {path}: layer names are offset; matched by creation order "
          f"({have_layers[0]} -> {want_layers[0]}, ...); shapes are checked when the weights are set")
    return out


def save_keras_weights(path, weights):
    """Write {name: array} in the Keras layout model_weights/<layer>/<layer>/<kernel|bias>:0 (+ layer_names /
    weight_names attributes), loadable by `load_keras_weights` and by Keras' `load_weights(by_name=True)`."""
    layers = []
    for n in weights:
        ly = n.split("/")[0]
        if ly not in layers:
            layers.append(ly)
    if os.path.exists(path):
        os.remove(path)
    with open_file(path, "w") as f:
        mw = f.create_group("model_weights")
        mw.attrs["layer_names"] = np.asarray([ly.encode() for ly in layers])
        mw.attrs["backend"] = np.asarray(b"sr4d")
        for ly in layers:
            g = mw.create_group(ly)
            names = [n for n in weights if n.split("/")[0] == ly]
            g.attrs["weight_names"] = np.asarray([f"{n}:0".encode() for n in names])
            gg = g.create_group(ly)
            for n in names:
                gg.create_dataset(n.split("/")[1] + ":0", data=np.asarray(weights[n], dtype=np.float32))
