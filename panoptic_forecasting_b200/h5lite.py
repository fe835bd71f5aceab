"""Minimal pure-Python HDF5 reader / writer for the depth containers on the input side of the bg path.

The reference reads its reprojected depth as `h5py.File(path)['<city>/<seq>/<frame:06d>/<start_fr>'][()]`, a
`[H, W, 3]` uint16 array per item (data/datasets/bg_dataset.py:184-196); the script that repacks the three
`..._depths.png` exports into that file is not in the reference tree.  h5py / libhdf5 are not
part of this image, so this module implements the subset of the HDF5 file format (HDF5 File Format Specification,
version 1/2 "classic" structures) those files use:

  reader  superblock v0/v1 (with user block / base address) and v2/v3; object headers v1 and v2 (incl. continuation
          blocks); old-style groups (symbol-table message: v1 B-tree + SNOD leaves + local heap) and compact new-style
          groups (link messages); dataspace v1/v2; fixed-point and IEEE float datatypes (little- or big-endian);
          data layout v3: compact, contiguous, chunked (v1 B-tree chunk index) with the deflate and shuffle filters.
          Dense-link groups (fractal heap), layout v4 and other filters raise NotImplementedError.
  writer  classic layout: superblock v0, symbol-table groups (one SNOD leaf per group, sized by the superblock's
          leaf K), object headers v1, contiguous little-endian datasets.  It is the repack step's output format.

Checked against a file written by libhdf5 itself where one is available (scipy ships a MATLAB 7.3 file:
tests/test_h5lite.py) and by round trips through the writer.
"""
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(IOError):
    pass


class _Buf:
    def __init__(self, data, base):
        self.d, self.base = data, base

    def at(self, addr, n):
        a = self.base + addr
        if a < 0 or a + n > len(self.d):
            raise H5Error("address 0x%x+%d outside the file" % (addr, n))
        return self.d[a:a + n]

    def u(self, addr, n):
        return int.from_bytes(self.at(addr, n), "little")


class Dataset:
    def __init__(self, f, shape, dtype, layout, filters):
        self._f, self.shape, self.dtype, self._layout, self._filters = f, shape, dtype, layout, filters

    def __getitem__(self, key):
        arr = self.read()
        return arr if key == () or key is Ellipsis else arr[key]

    def read(self):
        f, kind = self._f, self._layout[0]
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        nbytes = n * self.dtype.itemsize
        if kind == "compact":
            raw = self._layout[1][:nbytes]
        elif kind == "contiguous":
            addr = self._layout[1]
            raw = bytes(nbytes) if addr == UNDEF else f._b.at(addr, nbytes)
        else:
            return self._read_chunked()
        return np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape).copy()

    def _read_chunked(self):
        _, btree, cdims = self._layout                      # cdims excludes the trailing element-size entry
        out = np.zeros(self.shape, dtype=self.dtype)
        rank = len(self.shape)
        csize = int(np.prod(cdims)) * self.dtype.itemsize
        if btree == UNDEF:
            return out
        for offs, addr, size, mask in self._f._chunk_leaves(btree, rank):
            raw = self._f._b.at(addr, size)
            for i in reversed(range(len(self._filters))):   # undo the pipeline back to front
                fid, cvals = self._filters[i]
                if (mask >> i) & 1:
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cvals[0] if cvals else self.dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    k = len(a) // es
                    raw = a[:k * es].reshape(es, k).T.tobytes() + a[k * es:].tobytes()
                else:
                    raise NotImplementedError("HDF5 filter id %d" % fid)
            chunk = np.frombuffer(raw[:csize], dtype=self.dtype).reshape(cdims)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]
        return out


class Group:
    def __init__(self, f, links):
        self._f, self._links = f, links                      # name -> object header address

    def keys(self):
        return sorted(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._open(node._links[part])
        return node


class File(Group):
    """Read-only view of an HDF5 file: `File(path)['a/b/c'][()]` -> numpy array."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            data = fh.read()
        off = 0
        while data[off:off + 8] != SIG:                      # the superblock may follow a user block: 0, 512, 1024, ...
            off = 512 if off == 0 else off * 2
            if off + 8 > len(data):
                raise H5Error("not an HDF5 file: %s" % path)
        ver = data[off + 8]
        if ver in (0, 1):
            self._so, self._sl = data[off + 13], data[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            base = int.from_bytes(data[p:p + self._so], "little")
            p += 4 * self._so                                # base, free-space, end-of-file, driver-info addresses
            root_hdr = int.from_bytes(data[p + self._so:p + 2 * self._so], "little")   # symbol-table entry: link name offset, header address
        elif ver in (2, 3):
            self._so, self._sl = data[off + 9], data[off + 10]
            p = off + 12
            base = int.from_bytes(data[p:p + self._so], "little")
            root_hdr = int.from_bytes(data[p + 3 * self._so:p + 4 * self._so], "little")
        else:
            raise H5Error("unsupported superblock version %d" % ver)
        if self._so != 8 or self._sl != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        self._b = _Buf(data, base if base != 0 or ver >= 2 else off if off else 0)
        if base == 0 and off:                                # user block with relative addressing written as base 0
            self._b = _Buf(data, off)
        self._cache = {}
        root = self._open(root_hdr)
        if not isinstance(root, Group):
            raise H5Error("root object is not a group")
        Group.__init__(self, self, root._links)

    # ---- object headers -----------------------------------------------------------------------------------------
    def _messages(self, addr):
        b = self._b
        if b.at(addr, 4) == b"OHDR":
            flags = b.u(addr + 5, 1)
            p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            szl = 1 << (flags & 3)
            chunk = b.u(p, szl)
            blocks = [(p + szl, chunk)]
            order = bool(flags & 4)
            while blocks:
                q, n = blocks.pop(0)
                end = q + n
                while q + 4 <= end:
                    mtype, msize = b.u(q, 1), b.u(q + 1, 2)
                    q += 4 + (2 if order else 0)
                    body = b.at(q, msize)
                    if mtype == 0x10:
                        caddr, clen = struct.unpack("<QQ", body[:16])
                        blocks.append((caddr + 4, clen - 8))       # "OCHK" signature in front, checksum behind
                    elif mtype != 0:
                        yield mtype, body
                    q += msize
        else:
            if b.u(addr, 1) != 1:
                raise H5Error("bad object header at 0x%x" % addr)
            nmsg, hsize = b.u(addr + 2, 2), b.u(addr + 8, 4)
            blocks = [(addr + 16, hsize)]
            while blocks and nmsg > 0:
                q, n = blocks.pop(0)
                end = q + n
                while q + 8 <= end and nmsg > 0:
                    mtype, msize = b.u(q, 2), b.u(q + 2, 2)
                    body = b.at(q + 8, msize)
                    nmsg -= 1
                    if mtype == 0x10:
                        caddr, clen = struct.unpack("<QQ", body[:16])
                        blocks.append((caddr, clen))
                    elif mtype != 0:
                        yield mtype, body
                    q += 8 + msize

    def _open(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        shape = dtype = layout = None
        filters, links, symtab = [], {}, None
        for mtype, m in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(m)
            elif mtype == 0x03:
                dtype = self._datatype(m)
            elif mtype == 0x08:
                layout = self._layout_msg(m)
            elif mtype == 0x0B:
                filters = self._filters_msg(m)
            elif mtype == 0x11:
                symtab = struct.unpack("<QQ", m[:16])
            elif mtype == 0x06:
                name, target = self._link_msg(m)
                if target is not None:
                    links[name] = target
            elif mtype == 0x02:
                if len(m) >= 18 and struct.unpack("<Q", m[-16:-8])[0] != UNDEF:
                    raise NotImplementedError("dense link storage (fractal heap) is not supported")
        if layout is not None and shape is not None and dtype is not None:
            obj = Dataset(self, shape, dtype, layout, filters)
        else:
            if symtab is not None:
                links.update(self._symbol_table(*symtab))
            obj = Group(self, links)
        self._cache[addr] = obj
        return obj

    # ---- messages -----------------------------------------------------------------------------------------------
    def _dataspace(self, m):
        ver, rank, flags = m[0], m[1], m[2]
        p = 8 if ver == 1 else 4
        if ver == 2 and m[3] == 2:
            return ()                                        # null dataspace
        return tuple(int.from_bytes(m[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    def _datatype(self, m):
        cls, bits0, size = m[0] & 0x0F, m[1], struct.unpack("<I", m[4:8])[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits0 & 8 else "u", size))
        if cls == 1 and size in (2, 4, 8):
            return np.dtype("%sf%d" % (order, size))
        raise NotImplementedError("HDF5 datatype class %d size %d" % (cls, size))

    def _layout_msg(self, m):
        ver = m[0]
        if ver in (1, 2):                                    # libhdf5 <= 1.6: dimensionality, class, reserved, address, dims
            nd, cls = m[1], m[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack("<Q", m[p:p + 8])[0]
                p += 8
            dims = struct.unpack("<%dI" % nd, m[p:p + 4 * nd])
            p += 4 * nd
            if cls == 0:
                n = struct.unpack("<I", m[p:p + 4])[0]
                return ("compact", bytes(m[p + 4:p + 4 + n]))
            if cls == 1:
                return ("contiguous", addr)
            return ("chunked", addr, tuple(dims[:-1]))
        if ver != 3:
            raise NotImplementedError("data layout message version %d" % ver)
        cls = m[1]
        if cls == 0:
            n = struct.unpack("<H", m[2:4])[0]
            return ("compact", bytes(m[4:4 + n]))
        if cls == 1:
            return ("contiguous", struct.unpack("<Q", m[2:10])[0])
        if cls == 2:
            nd = m[2]
            addr = struct.unpack("<Q", m[3:11])[0]
            dims = struct.unpack("<%dI" % nd, m[11:11 + 4 * nd])
            return ("chunked", addr, tuple(dims[:-1]))
        raise NotImplementedError("data layout class %d" % cls)

    def _filters_msg(self, m):
        ver, n = m[0], m[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack("<H", m[p:p + 2])[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack("<H", m[p + 2:p + 4])[0]
                p += 4
            else:
                nlen = 0
                p += 2
            _, ncv = struct.unpack("<HH", m[p:p + 4])
            p += 4 + (((nlen + 7) // 8 * 8) if ver == 1 else nlen)
            cvals = struct.unpack("<%dI" % ncv, m[p:p + 4 * ncv])
            p += 4 * ncv + (4 if (ver == 1 and ncv % 2) else 0)
            out.append((fid, cvals))
        return out

    def _link_msg(self, m):
        flags = m[1]
        p = 2
        ltype = 0
        if flags & 8:
            ltype = m[p]; p += 1
        if flags & 4:
            p += 8
        if flags & 16:
            p += 1
        ls = 1 << (flags & 3)
        nlen = int.from_bytes(m[p:p + ls], "little"); p += ls
        name = m[p:p + nlen].decode("utf-8"); p += nlen
        return name, (struct.unpack("<Q", m[p:p + 8])[0] if ltype == 0 else None)

    # ---- old-style groups ---------------------------------------------------------------------------------------
    def _heap_string(self, heap_data_addr, off):
        b = self._b
        end = off
        while b.u(heap_data_addr + end, 1) != 0:
            end += 1
        return b.at(heap_data_addr + off, end - off).decode("utf-8")

    def _symbol_table(self, btree, heap):
        b = self._b
        if b.at(heap, 4) != b"HEAP":
            raise H5Error("bad local heap at 0x%x" % heap)
        heap_data = b.u(heap + 8 + 16, 8)
        links = {}
        stack = [btree]
        while stack:
            a = stack.pop()
            sig = b.at(a, 4)
            if sig == b"TREE":
                n = b.u(a + 6, 2)
                p = a + 8 + 16
                for i in range(n):
                    stack.append(b.u(p + 8 + 16 * i, 8))      # key, child, key, child, ..., key
            elif sig == b"SNOD":
                n = b.u(a + 6, 2)
                for i in range(n):
                    e = a + 8 + 40 * i
                    name_off, hdr = b.u(e, 8), b.u(e + 8, 8)
                    links[self._heap_string(heap_data, name_off)] = hdr
            else:
                raise H5Error("bad group node at 0x%x" % a)
        return links

    def _chunk_leaves(self, addr, rank):
        b = self._b
        if b.at(addr, 4) != b"TREE" or b.u(addr + 4, 1) != 1:
            raise H5Error("bad chunk B-tree at 0x%x" % addr)
        level, n = b.u(addr + 5, 1), b.u(addr + 6, 2)
        ksz = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(n):
            k = p + i * (ksz + 8)
            size, mask = b.u(k, 4), b.u(k + 4, 4)
            offs = tuple(b.u(k + 8 + 8 * j, 8) for j in range(rank))
            child = b.u(k + ksz, 8)
            if level == 0:
                yield offs, child, size, mask
            else:
                yield from self._chunk_leaves(child, rank)


# ----------------------------------------------------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------------------------------------------------
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 8 if dt.kind == "i" else 0
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # IEEE: bit fields = byte order LE, mantissa normalisation 2 (implied msb), sign location in bits 8-15
        if dt.itemsize == 4:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    raise NotImplementedError("dtype %s" % dt)


def _msg(mtype, body, flags=0):
    body = body + bytes((-len(body)) % 8)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def write(path, tree, chunks=None, deflate=False, shuffle=False):
    """tree: nested dict; leaves are numpy arrays (integer or float32/64).  Classic-format HDF5 file.
    chunks: optional chunk shape applied to every dataset of that rank (others stay contiguous); with it, `deflate` /
    `shuffle` add the two standard filters (one-level v1 B-tree chunk index)."""
    out = bytearray(96)                                      # superblock v0 (56 bytes + 40-byte root symbol-table entry)

    def alloc(data):
        while len(out) % 8:
            out.append(0)
        addr = len(out)
        out.extend(data)
        return addr

    max_entries = [1]

    def emit(node):
        if isinstance(node, dict):
            names = sorted(node)
            max_entries[0] = max(max_entries[0], len(names))
            children = [emit(node[k]) for k in names]
            heap = bytearray(8)                              # offset 0: the empty string (B-tree key 0)
            offs = []
            for k in names:
                offs.append(len(heap))
                kb = k.encode("utf-8") + b"\0"
                heap.extend(kb + bytes((-len(kb)) % 8))
            free_off = len(heap)
            heap.extend(struct.pack("<QQ", 1, 16))            # one free block closing the data segment: next = 1 (none), size
            heap_data = alloc(bytes(heap))
            heap_hdr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
            snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(names)))
            for o, c in zip(offs, children):
                snod.extend(struct.pack("<QQI4x16x", o, c, 0))
            snod_addr = alloc(bytes(snod)) if names else UNDEF
            tree_node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
            tree_node += struct.pack("<QQQ", 0, snod_addr, offs[-1] if names else 0) if names else struct.pack("<Q", 0)
            btree = alloc(tree_node)
            return alloc(_object_header([_msg(0x11, struct.pack("<QQ", btree, heap_hdr))]))
        arr = np.ascontiguousarray(node)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
        if chunks is not None and len(chunks) == arr.ndim and arr.ndim > 0:
            import itertools
            es = arr.dtype.itemsize
            entries = []
            for idx in itertools.product(*[range(0, s, c) for s, c in zip(arr.shape, chunks)]):
                blk = np.zeros(chunks, arr.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(idx, chunks, arr.shape))
                blk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
                raw = blk.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
                if deflate:
                    raw = zlib.compress(raw, 4)
                entries.append((idx, alloc(raw), len(raw)))
            if len(entries) > 65535:
                raise ValueError("too many chunks for a one-level index")
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF))
            for idx, addr, size in entries:
                node.extend(struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in idx) + struct.pack("<Q", 0))
                node.extend(struct.pack("<Q", addr))
            node.extend(struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0))
            btree = alloc(bytes(node))
            layout = struct.pack("<BBBQ", 3, 2, arr.ndim + 1, btree) + struct.pack("<%dI" % (arr.ndim + 1), *chunks, es)
            msgs = [_msg(0x01, space), _msg(0x03, _dtype_msg(arr.dtype), 1), _msg(0x08, layout)]
            filt = b""
            nf = 0
            if shuffle:
                filt += struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", es) + bytes(4); nf += 1
            if deflate:
                filt += struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 4) + bytes(4); nf += 1
            if nf:
                msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, nf) + filt))
            return alloc(_object_header(msgs))
        data = alloc(arr.tobytes())
        layout = struct.pack("<BBQQ", 3, 1, data, arr.nbytes)
        return alloc(_object_header([_msg(0x01, space), _msg(0x03, _dtype_msg(arr.dtype), 1), _msg(0x08, layout)]))

    root = emit(tree)
    leaf_k = max(4, (max_entries[0] + 1) // 2)               # a SNOD holds up to 2K entries
    if leaf_k > 32767:
        raise ValueError("too many entries in one group")
    sb = SIG + struct.pack("<BBBBBBBxHHI", 0, 0, 0, 0, 0, 8, 8, leaf_k, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF)
    sb += struct.pack("<QQI4x16x", 0, root, 0)
    out[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(out))


def read_bg_depth(path, city, seq, frame, start_fr):
    """The lookup BGDataset does (bg_dataset.py:186-187,196): `[H, W, 3]` uint16, one plane per input frame."""
    return File(path)["%s/%s/%06d/%s" % (city, seq, frame, start_fr)][()]
