"""Minimal pure-Python HDF5 reader / writer for Keras weight files (no h5py in this image).

Covers exactly what `Model.save_weights` / `Model.save` of Keras-2.0.8 write through h5py
(Keras-2.0.8/keras/engine/topology.py:2555-2590, 2845-2872) and what its loaders read (:2590-2630, 3107-3330):

  file      superblock version 0 (or 1), 8-byte offsets / lengths
  groups    "old style": object header v1 + Symbol Table message -> v1 B-tree ("TREE") of symbol nodes ("SNOD") +
            local heap ("HEAP") holding the link names
  datasets  contiguous or compact layout (Keras calls create_dataset(name, shape, dtype) without chunking),
            little-endian IEEE floats / integers
  attrs     v1-v3 attribute messages in the object header (+ continuation blocks): numeric scalars / arrays,
            fixed-length strings (numpy 'S' arrays: `layer_names`, `weight_names`), variable-length strings through the
            global heap ("GCOL"; what a Python-3 h5py writes for str attributes)

Not covered (raises H5Error): chunked / compressed datasets, new-style (link-message / fractal-heap) groups,
object header v2.  Byte layouts follow the HDF5 File Format Specification, version 2.0, sections III.A-E and IV.A.
The writer emits files in the same subset, so this reader -- which is also checked against a file written by the real
library, Keras-2.0.8/examples/mymodel.h5 (tests/test_h5lite.py) -- is its first consumer.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ============================================================================================ reader
class _Datatype(object):
    def __init__(self, cls, size, np_dtype=None, strpad=None, vlen_str=False, base=None):
        self.cls, self.size, self.np_dtype, self.strpad, self.vlen_str, self.base = cls, size, np_dtype, strpad, vlen_str, base


def _parse_datatype(buf, off):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls, p = cv & 0x0F, off + 8
    if cls == 0:                                           # fixed point
        order = ">" if (b0 & 1) else "<"
        signed = bool(b0 & 0x08)
        return _Datatype(cls, size, np.dtype("%s%s%d" % (order, "i" if signed else "u", size)))
    if cls == 1:                                           # floating point
        order = ">" if (b0 & 1) else "<"
        if size not in (2, 4, 8):
            raise H5Error("unsupported float size %d" % size)
        return _Datatype(cls, size, np.dtype("%sf%d" % (order, size)))
    if cls == 3:                                           # fixed-length string
        return _Datatype(cls, size, np.dtype("S%d" % size), strpad=b0 & 0x0F)
    if cls == 9:                                           # variable length: type in bits 0-3 of b0 (1 = string)
        base = _parse_datatype(buf, p)
        return _Datatype(cls, size, vlen_str=(b0 & 0x0F) == 1, base=base)
    raise H5Error("unsupported datatype class %d" % cls)


def _parse_dataspace(buf, off):
    ver = buf[off]
    if ver == 1:
        rank, flags = buf[off + 1], buf[off + 2]
        p = off + 8
    elif ver == 2:
        rank, flags, typ = buf[off + 1], buf[off + 2], buf[off + 3]
        p = off + 4
        if typ == 2:                                       # null dataspace
            return None
    else:
        raise H5Error("unsupported dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, p)) if rank else ()


class _Node(object):
    """An object (group or dataset) located by its object-header address."""

    def __init__(self, f, addr, name="/"):
        self.file, self.addr, self.name = f, addr, name
        self._msgs = None
        self._attrs = None
        self._links = None

    # -- object header v1 ------------------------------------------------------------------------
    def _messages(self):
        if self._msgs is not None:
            return self._msgs
        buf = self.file.buf
        a = self.addr
        if buf[a:a + 4] == b"OHDR":
            raise H5Error("object header version 2 (libver='latest' files) is not supported")
        ver, _, nmsg, _ref, hsize = struct.unpack_from("<BBHII", buf, a)
        if ver != 1:
            raise H5Error("unsupported object header version %d at %#x" % (ver, a))
        blocks = [(a + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, p)
                body = p + 8
                if mtype == 0x0010:                        # continuation
                    off, ln = struct.unpack_from("<QQ", buf, body)
                    blocks.append((off, ln))
                msgs.append((mtype, body, msize, mflags))
                p = body + msize
        self._msgs = msgs
        return msgs

    def _find(self, mtype):
        for t, body, size, _ in self._messages():
            if t == mtype:
                return body, size
        return None

    # -- attributes --------------------------------------------------------------------------------
    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for t, body, size, flags in self._messages():
                if t == 0x000C:
                    if flags & 0x02:
                        raise H5Error("shared attribute messages are not supported")
                    k, v = self._parse_attr(body)
                    self._attrs[k] = v
        return self._attrs

    def _parse_attr(self, p):
        buf = self.file.buf
        ver = buf[p]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", buf, p + 2)
            q = p + 8
            name = bytes(buf[q:q + nsz]).split(b"\0")[0].decode("utf8"); q += _pad8(nsz)
            dt = _parse_datatype(buf, q); q += _pad8(tsz)
            shape = _parse_dataspace(buf, q); q += _pad8(ssz)
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", buf, p + 2)
            q = p + 8 + (1 if ver == 3 else 0)
            name = bytes(buf[q:q + nsz]).split(b"\0")[0].decode("utf8"); q += nsz
            dt = _parse_datatype(buf, q); q += tsz
            shape = _parse_dataspace(buf, q); q += ssz
        else:
            raise H5Error("unsupported attribute message version %d" % ver)
        return name, self.file._read_values(dt, shape, q)

    # -- group ---------------------------------------------------------------------------------------
    def _load_links(self):
        if self._links is not None:
            return self._links
        st = self._find(0x0011)
        if st is None:
            if self._find(0x0002) is not None or self._find(0x0006) is not None:
                raise H5Error("new-style (link message) groups are not supported: %s" % self.name)
            self._links = None
            return None
        buf = self.file.buf
        btree, heap = struct.unpack_from("<QQ", buf, st[0])
        if bytes(buf[heap:heap + 4]) != b"HEAP":
            raise H5Error("bad local heap signature at %#x" % heap)
        _dsize, _free, daddr = struct.unpack_from("<QQQ", buf, heap + 8)
        links = {}

        def name_at(off):
            e = buf.find(b"\0", daddr + off) if isinstance(buf, (bytes, bytearray)) else bytes(buf[daddr + off:daddr + off + 512]).find(b"\0") + daddr + off
            return bytes(buf[daddr + off:e]).decode("utf8")

        def walk(addr):
            sig = bytes(buf[addr:addr + 4])
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", buf, addr + 4)
                if ntype != 0:
                    raise H5Error("unexpected B-tree node type %d in a group" % ntype)
                p = addr + 8 + 16                          # past the sibling pointers
                for i in range(used):
                    child = struct.unpack_from("<Q", buf, p + 8 + i * 16)[0]     # key_i (8), child_i (8), ...
                    walk(child)
            elif sig == b"SNOD":
                n = struct.unpack_from("<H", buf, addr + 6)[0]
                p = addr + 8
                for i in range(n):
                    noff, oaddr = struct.unpack_from("<QQ", buf, p + i * 40)
                    links[name_at(noff)] = oaddr
            else:
                raise H5Error("bad group node signature %r at %#x" % (sig, addr))

        walk(btree)
        self._links = links
        return links

    @property
    def is_group(self):
        return self._load_links() is not None

    def keys(self):
        """Link names in the order h5py returns them for old-style groups: sorted by name."""
        links = self._load_links()
        if links is None:
            raise H5Error("%s is a dataset, not a group" % self.name)
        return sorted(links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, key):
        node = self
        for part in [s for s in key.split("/") if s]:
            links = node._load_links()
            if links is None or part not in links:
                raise KeyError("%s not found in %s" % (key, self.name))
            node = _Node(self.file, links[part], (node.name.rstrip("/") + "/" + part))
        return node

    # -- dataset -------------------------------------------------------------------------------------
    @property
    def shape(self):
        s = self._find(0x0001)
        return _parse_dataspace(self.file.buf, s[0]) if s else None

    def value(self):
        buf = self.file.buf
        t, s, l = self._find(0x0003), self._find(0x0001), self._find(0x0008)
        if not (t and s and l):
            raise H5Error("%s is not a dataset" % self.name)
        dt = _parse_datatype(buf, t[0])
        shape = _parse_dataspace(buf, s[0])
        p = l[0]
        ver = buf[p]
        if ver == 3:
            cls = buf[p + 1]
            if cls == 1:
                addr, _size = struct.unpack_from("<QQ", buf, p + 2)
            elif cls == 0:
                addr = p + 4
            else:
                raise H5Error("chunked dataset %s: not supported (Keras writes contiguous datasets)" % self.name)
        elif ver in (1, 2):
            rank, cls = buf[p + 1], buf[p + 2]
            if cls == 1:
                addr = struct.unpack_from("<Q", buf, p + 8)[0]
            elif cls == 0:
                addr = p + 8 + 4 * rank + 4
            else:
                raise H5Error("chunked dataset %s: not supported" % self.name)
        else:
            raise H5Error("unsupported data layout version %d" % ver)
        if addr == UNDEF:
            n = int(np.prod(shape)) if shape else 1
            return np.zeros(shape, dt.np_dtype) if n else np.zeros(shape, dt.np_dtype)
        return self.file._read_values(dt, shape, addr)

    def __array__(self, dtype=None, copy=None):
        v = self.value()
        return v.astype(dtype) if dtype is not None else v


class File(_Node):
    """Read-only view of an HDF5 file: File(path)['group/dataset'].value(), .attrs, .keys()."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        buf = self.buf
        base = 0
        while buf[base:base + 8] != SIG:                   # the superblock may sit at 0, 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base >= len(buf):
                raise H5Error("%s: not an HDF5 file" % path)
        ver = buf[base + 8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d is not supported (only the classic 0/1 layout)" % ver)
        so, sl = buf[base + 13], buf[base + 14]
        if (so, sl) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = base + 24 + (4 if ver == 1 else 0)
        _base_addr, _fs, _eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        p += 32
        _noff, root = struct.unpack_from("<QQ", buf, p)
        _Node.__init__(self, self, root, "/")

    def close(self):
        pass

    def _read_values(self, dt, shape, addr):
        buf = self.buf
        n = int(np.prod(shape)) if shape else 1
        if shape is None:
            return None
        if dt.cls == 9:
            if not dt.vlen_str:
                raise H5Error("variable-length sequences other than strings are not supported")
            out = []
            for i in range(n):
                ln, gaddr, idx = struct.unpack_from("<IQI", buf, addr + i * 16)
                out.append(self._global_heap_object(gaddr, idx)[:ln])
            arr = np.array(out, dtype=object)
            return arr.reshape(shape) if shape else out[0]
        a = np.frombuffer(buf, dtype=dt.np_dtype, count=n, offset=addr)
        if dt.np_dtype.byteorder == ">":
            a = a.astype(dt.np_dtype.newbyteorder("<"))
        a = a.reshape(shape).copy() if shape else a[0]
        return a

    def _global_heap_object(self, addr, idx):
        buf = self.buf
        if bytes(buf[addr:addr + 4]) != b"GCOL":
            raise H5Error("bad global heap signature at %#x" % addr)
        size = struct.unpack_from("<Q", buf, addr + 8)[0]
        p, end = addr + 16, addr + size
        while p + 16 <= end:
            i, _ref, _, osz = struct.unpack_from("<HHIQ", buf, p)
            if i == idx:
                return bytes(buf[p + 16:p + 16 + osz])
            if i == 0:
                break
            p += 16 + _pad8(osz)
        raise H5Error("global heap object %d not found" % idx)


# ============================================================================================ writer
def _dt_message(arr):
    """Datatype message body for a numpy array (floats, ints, fixed-length byte strings)."""
    dt = arr.dtype
    if dt.kind == "f":
        size = dt.itemsize
        exp, man, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        # class 1 version 1; bits: little-endian, implied msb-set normalisation (2 << 4), sign at the top bit
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
        return head + struct.pack("<HHBBBBI", 0, size * 8, man, exp, 0, man, bias)
    if dt.kind in "iu":
        size = dt.itemsize
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, size)
        return head + struct.pack("<HH", 0, size * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)        # null-padded (numpy 'S'), ASCII
    raise H5Error("cannot write dtype %s" % dt)


def _ds_message(shape):
    rank = len(shape)
    body = struct.pack("<BBBBI", 1, rank, 0, 0, 0)
    return body + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _msg(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _attr_message(name, value):
    if isinstance(value, (bytes, str)):
        value = np.array(value.encode("utf8") if isinstance(value, str) else value)
    arr = np.asarray(value)
    if arr.dtype.kind == "U":
        arr = np.char.encode(arr, "utf8")
    if arr.dtype.kind == "O":
        arr = np.array([x if isinstance(x, bytes) else str(x).encode("utf8") for x in arr.ravel()]).reshape(arr.shape)
    if arr.dtype.kind == "S" and arr.dtype.itemsize == 0:
        arr = arr.astype("S1")
    nm = name.encode("utf8") + b"\0"
    dt, ds = _dt_message(arr), _ds_message(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
    body += dt + b"\0" * (_pad8(len(dt)) - len(dt))
    body += ds + b"\0" * (_pad8(len(ds)) - len(ds))
    body += np.ascontiguousarray(arr).tobytes()
    if len(body) > 65000:
        raise H5Error("attribute %s is %d bytes: object-header attributes hold at most 64 KiB" % (name, len(body)))
    return _msg(0x000C, body)


class Writer(object):
    """Builds the tree in memory, lays the file out on close():  w = Writer(path); g = w.root.group('conv1');
    g.attrs['weight_names'] = [...]; g.dataset('conv1/kernel:0', array); w.close()."""

    LEAF_K, INTERNAL_K = 4, 16           # the library's default group B-tree sizing (8 links per symbol node, 32 children per node)

    class Group(object):
        def __init__(self):
            self.attrs, self.children = {}, {}

        def group(self, name):
            node = self
            for part in [s for s in name.split("/") if s]:
                nxt = node.children.get(part)
                if nxt is None:
                    nxt = node.children[part] = Writer.Group()
                if not isinstance(nxt, Writer.Group):
                    raise H5Error("%s is a dataset" % part)
                node = nxt
            return node

        create_group = group

        def dataset(self, name, value):
            parts = [s for s in name.split("/") if s]
            g = self.group("/".join(parts[:-1])) if len(parts) > 1 else self
            d = Writer.Dataset(np.array(value, order="C"))        # (ascontiguousarray would turn a scalar into shape (1,))
            g.children[parts[-1]] = d
            return d

    class Dataset(object):
        def __init__(self, value):
            self.attrs, self.value = {}, value

    def __init__(self, path):
        self.path = path
        self.root = Writer.Group()
        self.chunks = []
        self.pos = 0

    # -- raw allocation
    def _alloc(self, data, align=8):
        pad = (-self.pos) % align
        if pad:
            self.chunks.append(b"\0" * pad)
            self.pos += pad
        addr = self.pos
        self.chunks.append(data)
        self.pos += len(data)
        return addr

    def _object_header(self, msgs):
        body = b"".join(msgs)
        hdr = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4
        return self._alloc(hdr + body)

    def _write_dataset(self, d):
        arr = d.value
        if arr.dtype.kind == "f" and arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        daddr = self._alloc(raw) if raw else UNDEF
        msgs = [_msg(0x0001, _ds_message(arr.shape)), _msg(0x0003, _dt_message(arr), flags=1),
                _msg(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),    # fill value v2 as h5py writes it: late / if-set / default
                _msg(0x0008, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]
        msgs += [_attr_message(k, v) for k, v in d.attrs.items()]
        return self._object_header(msgs)

    def _write_group(self, g):
        # children first (their header addresses go into the symbol nodes)
        names = sorted(g.children, key=lambda s: s.encode("utf8"))
        addrs = {}
        for n in names:
            c = g.children[n]
            addrs[n] = self._write_group(c) if isinstance(c, Writer.Group) else self._write_dataset(c)
        # local heap: offset 0 holds the empty string (the B-tree's first key)
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            b = n.encode("utf8") + b"\0"
            heap += b + b"\0" * (_pad8(len(b)) - len(b))
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)                  # one free block: (next = 1: none, size 16) -- libhdf5 wants >= 16 free bytes layout-valid
        hdata = self._alloc(bytes(heap))
        haddr = self._alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free_off, hdata))
        # symbol nodes of <= 2*LEAF_K links, then B-tree levels of <= 2*INTERNAL_K children until one root node remains
        per = 2 * self.LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        nodes = []                                         # (address, heap offset of the greatest name below)
        for chunk in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                body += struct.pack("<QQII", offs[n], addrs[n], 0, 0) + b"\0" * 16
            body += b"\0" * (40 * (per - len(chunk)))
            nodes.append((self._alloc(body), offs[chunk[-1]] if chunk else 0))
        fan = 2 * self.INTERNAL_K
        node_size = 24 + 8 * (fan + 1) + 8 * fan
        level = 0
        while True:
            groups = [nodes[i:i + fan] for i in range(0, len(nodes), fan)]
            start = self._alloc(b"")                       # 8-byte aligned position of this level's first node
            out, first_key = [], 0
            for gi, grp in enumerate(groups):
                left = start + (gi - 1) * node_size if gi > 0 else UNDEF
                right = start + (gi + 1) * node_size if gi + 1 < len(groups) else UNDEF
                t = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right) + struct.pack("<Q", first_key)
                for a, k in grp:
                    t += struct.pack("<QQ", a, k)
                t += b"\0" * (node_size - len(t))
                addr = self._alloc(t)
                assert addr == start + gi * node_size
                first_key = grp[-1][1]
                out.append((addr, grp[-1][1]))
            nodes = out
            level += 1
            if len(nodes) == 1:
                break
        baddr = nodes[0][0]
        msgs = [_msg(0x0011, struct.pack("<QQ", baddr, haddr))]
        msgs += [_attr_message(k, v) for k, v in g.attrs.items()]
        oaddr = self._object_header(msgs)
        g._btree, g._heap = baddr, haddr
        return oaddr

    def close(self):
        self.chunks, self.pos = [], 0
        self._alloc(b"\0" * 96)                            # superblock placeholder
        root = self._write_group(self.root)
        eof = self.pos
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", self.root._btree, self.root._heap)
        assert len(sb) == 96, len(sb)
        data = b"".join(self.chunks)
        with open(self.path, "wb") as fh:
            fh.write(sb + data[96:])
