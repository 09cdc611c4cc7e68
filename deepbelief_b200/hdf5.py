"""Minimal HDF5 reader/writer for the dataset files the reference ships (h5py is not available).

The reference loads `datapoints` / `labels` with h5py's in-memory core driver (data.py:51-57). Its
files are all of one simple shape: superblock version 0, a root group stored as a symbol table
(B-tree 'TREE' + local heap 'HEAP' + one 'SNOD' leaf), version-1 object headers, and contiguous
little-endian IEEE datasets (data layout message version 3, class 1). This module parses exactly that
subset — by walking superblock -> root symbol table -> object headers, not by hard-coded offsets —
and can write the same subset, so synthetic datasets can be produced for the examples and tests.
"""
import struct

import numpy as np

_SIGNATURE = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    pass


def _u(fmt, buf, off):
    return struct.unpack_from('<' + fmt, buf, off)


class File:
    """Read-only view: `File(path)['datapoints']` -> numpy array (a copy held in memory)."""

    def __init__(self, path):
        with open(path, 'rb') as f:
            self._buf = f.read()
        b = self._buf
        if b[:8] != _SIGNATURE:
            raise HDF5FormatError('%s: not an HDF5 file' % path)
        if b[8] != 0:
            raise HDF5FormatError('%s: superblock version %d is not supported (only 0)' % (path, b[8]))
        if b[13] != 8 or b[14] != 8:
            raise HDF5FormatError('only 8-byte offsets/lengths are supported')
        self._base = _u('Q', b, 24)[0]
        # root group symbol table entry at byte 56: link name offset, object header address, cache type, scratch
        _, root_ohdr, cache_type = _u('QQI', b, 56)
        if cache_type == 1:
            btree, heap = _u('QQ', b, 80)
        else:
            btree, heap = self._symbol_table_message(root_ohdr)
        self._datasets = {}
        for name, ohdr in self._iter_group(btree, heap):
            self._datasets[name] = ohdr

    # ---- group traversal ------------------------------------------------------------------------------
    def _symbol_table_message(self, ohdr):
        for mtype, off, size in self._messages(ohdr):
            if mtype == 0x11:
                return _u('QQ', self._buf, off)
        raise HDF5FormatError('group object header has no symbol table message')

    def _iter_group(self, btree, heap):
        b = self._buf
        if b[heap:heap + 4] != b'HEAP':
            raise HDF5FormatError('bad local heap signature')
        heap_data = _u('Q', b, heap + 24)[0]

        def walk(node):
            if b[node:node + 4] != b'TREE':
                raise HDF5FormatError('bad B-tree signature')
            node_type, level, nent = _u('BBH', b, node + 4)
            if node_type != 0:
                raise HDF5FormatError('unexpected B-tree node type %d' % node_type)
            p = node + 24 + 8          # skip sibling pointers and key 0
            for _ in range(nent):
                child = _u('Q', b, p)[0]
                p += 16                # child pointer + next key
                if level > 0:
                    for item in walk(child):
                        yield item
                else:
                    if b[child:child + 4] != b'SNOD':
                        raise HDF5FormatError('bad symbol table node signature')
                    nsym = _u('H', b, child + 6)[0]
                    q = child + 8
                    for _ in range(nsym):
                        name_off, ohdr = _u('QQ', b, q)
                        start = heap_data + name_off
                        end = b.index(b'\0', start)
                        yield b[start:end].decode('ascii'), ohdr
                        q += 40
        return list(walk(btree))

    # ---- object headers -----------------------------------------------------------------------------------
    def _messages(self, ohdr):
        b = self._buf
        version, _, nmsg, _, hsize = _u('BBHII', b, ohdr)
        if version != 1:
            raise HDF5FormatError('object header version %d is not supported (only 1)' % version)
        out = []
        blocks = [(ohdr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = _u('HHB', b, p)
                body = p + 8
                if mtype == 0x10:      # continuation
                    coff, clen = _u('QQ', b, body)
                    blocks.append((self._base + coff, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def _read_dataset(self, ohdr):
        b = self._buf
        shape = dtype = addr = nbytes = None
        for mtype, off, size in self._messages(ohdr):
            if mtype == 0x1:           # dataspace
                ver, rank, flags = _u('BBB', b, off)
                if ver != 1:
                    raise HDF5FormatError('dataspace version %d is not supported' % ver)
                shape = _u('%dQ' % rank, b, off + 8)
            elif mtype == 0x3:         # datatype
                cls_ver = b[off]
                cls, bits0 = cls_ver & 0x0F, b[off + 1]
                tsize = _u('I', b, off + 4)[0]
                if bits0 & 1:
                    raise HDF5FormatError('big-endian datasets are not supported')
                if cls == 1:
                    dtype = {4: np.float32, 8: np.float64}.get(tsize)
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = {(1, True): np.int8, (2, True): np.int16, (4, True): np.int32, (8, True): np.int64,
                             (1, False): np.uint8, (2, False): np.uint16, (4, False): np.uint32,
                             (8, False): np.uint64}.get((tsize, signed))
                if dtype is None:
                    raise HDF5FormatError('unsupported datatype class %d size %d' % (cls, tsize))
            elif mtype == 0x8:         # data layout
                ver, lclass = _u('BB', b, off)
                if ver != 3 or lclass != 1:
                    raise HDF5FormatError('only contiguous layout (version 3, class 1) is supported')
                addr, nbytes = _u('QQ', b, off + 2)
        if shape is None or dtype is None or addr is None:
            raise HDF5FormatError('dataset object header lacks dataspace / datatype / layout')
        count = int(np.prod(shape)) if shape else 1
        if addr == _UNDEF:
            return np.zeros(shape, dtype=dtype)
        arr = np.frombuffer(b, dtype=np.dtype(dtype).newbyteorder('<'), count=count, offset=self._base + addr)
        return arr.reshape(shape).astype(dtype, copy=True)

    # ---- mapping interface -----------------------------------------------------------------------------------
    def keys(self):
        return list(self._datasets.keys())

    def __contains__(self, name):
        return name in self._datasets

    def __getitem__(self, name):
        if name not in self._datasets:
            raise KeyError(name)
        return self._read_dataset(self._datasets[name])


def write(path, datasets):
    """Write {name: 2-D float32/float64 array} in the layout the reference's files use."""
    names = sorted(datasets.keys())          # symbol table nodes are ordered by name
    if not 1 <= len(names) <= 8:
        raise ValueError('between 1 and 8 datasets supported')
    arrays = []
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        arrays.append(a)

    def pad8(x):
        return (x + 7) & ~7

    # local heap data: empty string at 0, then the names, 8-byte aligned
    heap_items = [b'\0' * 8]
    name_off = []
    pos = 8
    for n in names:
        raw = n.encode('ascii') + b'\0'
        raw += b'\0' * (pad8(len(raw)) - len(raw))
        name_off.append(pos)
        heap_items.append(raw)
        pos += len(raw)
    heap_data = b''.join(heap_items)
    heap_data_size = max(pad8(len(heap_data)) + 16, 88)
    heap_data += b'\0' * (heap_data_size - len(heap_data))
    free_off = pad8(pos)

    root_ohdr = 96
    btree = 136
    btree_size = 24 + (2 * 16 + 1) * 16 + 8    # generous: node with K=16
    heap = pad8(btree + btree_size)
    heap_data_addr = heap + 32
    snod = pad8(heap_data_addr + heap_data_size)
    snod_size = 8 + 2 * 4 * 40                  # leaf K = 4 -> 8 entries
    ohdr0 = pad8(snod + snod_size)
    ohdr_size = 16 + 256
    data0 = pad8(ohdr0 + ohdr_size * len(names))
    data_addr = []
    p = data0
    for a in arrays:
        data_addr.append(p)
        p = pad8(p + a.nbytes)
    eof = p

    out = bytearray(eof)
    out[0:8] = _SIGNATURE
    out[8:16] = bytes([0, 0, 0, 0, 0, 8, 8, 0])
    struct.pack_into('<HHI', out, 16, 4, 16, 0)
    struct.pack_into('<QQQQ', out, 24, 0, _UNDEF, eof, _UNDEF)
    struct.pack_into('<QQII', out, 56, 0, root_ohdr, 1, 0)
    struct.pack_into('<QQ', out, 80, btree, heap)
    # root object header: one symbol-table message
    struct.pack_into('<BBHII', out, root_ohdr, 1, 0, 1, 1, 24)
    struct.pack_into('<HHBBBB', out, root_ohdr + 16, 0x11, 16, 0, 0, 0, 0)
    struct.pack_into('<QQ', out, root_ohdr + 24, btree, heap)
    # B-tree: one leaf child
    out[btree:btree + 4] = b'TREE'
    struct.pack_into('<BBH', out, btree + 4, 0, 0, 1)
    struct.pack_into('<QQ', out, btree + 8, _UNDEF, _UNDEF)
    struct.pack_into('<QQQ', out, btree + 24, 0, snod, name_off[-1])
    # local heap
    out[heap:heap + 4] = b'HEAP'
    struct.pack_into('<QQQ', out, heap + 8, heap_data_size, free_off, heap_data_addr)
    out[heap_data_addr:heap_data_addr + heap_data_size] = heap_data
    if free_off + 16 <= heap_data_size:
        struct.pack_into('<QQ', out, heap_data_addr + free_off, 1, heap_data_size - free_off)
    # symbol table node
    out[snod:snod + 4] = b'SNOD'
    struct.pack_into('<BBH', out, snod + 4, 1, 0, len(names))
    for i, (n, a) in enumerate(zip(names, arrays)):
        oh = ohdr0 + i * ohdr_size
        struct.pack_into('<QQII', out, snod + 8 + 40 * i, name_off[i], oh, 0, 0)
        struct.pack_into('<BBHII', out, oh, 1, 0, 4, 1, 256)
        q = oh + 16
        # dataspace v1 with max dims
        rank = a.ndim
        body = struct.pack('<BBBBI', 1, rank, 1, 0, 0) + struct.pack('<%dQ' % rank, *a.shape) * 2
        struct.pack_into('<HHBBBB', out, q, 0x1, len(body), 0, 0, 0, 0)
        out[q + 8:q + 8 + len(body)] = body
        q += 8 + len(body)
        # datatype: IEEE little-endian float
        if a.dtype == np.float32:
            body = bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack('<I', 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
        else:
            body = bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack('<I', 8) + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
        body += b'\0' * (pad8(len(body)) - len(body))
        struct.pack_into('<HHBBBB', out, q, 0x3, len(body), 1, 0, 0, 0)
        out[q + 8:q + 8 + len(body)] = body
        q += 8 + len(body)
        # fill value (v2, never written)
        body = bytes([2, 2, 2, 1, 0, 0, 0, 0])
        struct.pack_into('<HHBBBB', out, q, 0x5, len(body), 1, 0, 0, 0)
        out[q + 8:q + 8 + len(body)] = body
        q += 8 + len(body)
        # layout v3 contiguous
        body = struct.pack('<BBQQ', 3, 1, data_addr[i], a.nbytes) + b'\0' * 6
        struct.pack_into('<HHBBBB', out, q, 0x8, len(body), 1, 0, 0, 0)
        out[q + 8:q + 8 + len(body)] = body
        q += 8 + len(body)
        # NIL message fills the rest of the header block
        rest = oh + 16 + 256 - q - 8
        struct.pack_into('<HHBBBB', out, q, 0x0, rest, 0, 0, 0, 0)
        # header message count includes the NIL message
        struct.pack_into('<H', out, oh + 2, 5)
        out[data_addr[i]:data_addr[i] + a.nbytes] = a.astype(a.dtype.newbyteorder('<')).tobytes()
    with open(path, 'wb') as f:
        f.write(bytes(out))
