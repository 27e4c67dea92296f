"""A minimal reader for the subset of HDF5 the reference's MT-CKD table uses
(pyLBL/mt_ckd/mt-ckd.nc, a netCDF-4 file): superblock version 0, version-2 object headers with
continuation chunks, links stored densely in a fractal heap (direct blocks scanned for link
messages), contiguous / compact / single-chunk float datasets without filters, and compact
attributes.  Neither h5py nor netCDF4 exists in the build image; this is only used once, by
tools/convert_mt_ckd.py.  It is not a general HDF5 library and says so when it meets
something it does not know.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Unsupported(Exception):
    pass


def _messages(buf, start, end, creation_order):
    """Yields (type, flags, payload offset, payload size) of the messages in buf[start:end]."""
    pos = start
    while pos + 4 <= end:
        mtype = buf[pos]
        size, = struct.unpack_from("<H", buf, pos + 1)
        flags = buf[pos + 3]
        pos += 4
        if creation_order:
            pos += 2
        if pos + size > end:
            break
        yield mtype, flags, pos, size
        pos += size


class ObjectHeader(object):
    def __init__(self, buf, address):
        if buf[address:address + 4] != b"OHDR":
            raise Unsupported(f"no version-2 object header at {address}")
        if buf[address + 4] != 2:
            raise Unsupported("object header version")
        flags = buf[address + 5]
        pos = address + 6
        if flags & 0x20:
            pos += 16
        if flags & 0x10:
            pos += 4
        width = 1 << (flags & 0x3)
        chunk0 = int.from_bytes(buf[pos:pos + width], "little")
        pos += width
        self.messages = []
        creation_order = bool(flags & 0x04)
        blocks = [(pos, pos + chunk0)]
        while blocks:
            start, end = blocks.pop(0)
            for mtype, mflags, off, size in _messages(buf, start, end, creation_order):
                if mtype == 0x10:      # continuation
                    caddr, clen = struct.unpack_from("<QQ", buf, off)
                    if buf[caddr:caddr + 4] != b"OCHK":
                        raise Unsupported("continuation block signature")
                    blocks.append((caddr + 4, caddr + clen - 4))
                else:
                    self.messages.append((mtype, mflags, off, size))


def _datatype(buf, off):
    cls = buf[off] & 0x0F
    bits0 = buf[off + 1]
    size, = struct.unpack_from("<I", buf, off + 4)
    if cls == 1 and not (bits0 & 1):       # little-endian IEEE float
        return {4: np.dtype("<f4"), 8: np.dtype("<f8")}[size], 8 + 12
    if cls == 0 and not (bits0 & 1):       # little-endian integer
        signed = bool(bits0 & 0x08)
        return np.dtype(("<i" if signed else "<u") + str(size)), 8 + 4
    if cls == 3:                           # fixed-length string
        return np.dtype(f"S{size}"), 8
    return None, None


def _dataspace(buf, off):
    version, rank, flags = buf[off], buf[off + 1], buf[off + 2]
    if version == 1:
        pos = off + 8
    elif version == 2:
        pos = off + 4
    else:
        raise Unsupported("dataspace version")
    dims = struct.unpack_from(f"<{rank}Q", buf, pos) if rank else ()
    size = (8 if version == 1 else 4) + rank * 8 * (2 if flags & 1 else 1)
    return tuple(int(x) for x in dims), size


def _attribute(buf, off):
    version = buf[off]
    if version == 3:
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, off + 2)
        pos = off + 9
        pad = lambda n: n
    elif version == 1:
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, off + 2)
        pos = off + 8
        pad = lambda n: (n + 7) & ~7
    elif version == 2:
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", buf, off + 2)
        pos = off + 8
        pad = lambda n: n
    else:
        raise Unsupported("attribute version")
    name = buf[pos:pos + name_size].split(b"\0")[0].decode()
    pos += pad(name_size)
    dtype, _ = _datatype(buf, pos)
    pos += pad(dt_size)
    dims, _ = _dataspace(buf, pos)
    pos += pad(ds_size)
    if dtype is None:
        return name, None
    count = int(np.prod(dims)) if dims else 1
    value = np.frombuffer(buf, dtype=dtype, count=count, offset=pos)
    return name, (value[0] if not dims else value.copy())


class Dataset(object):
    def __init__(self, buf, address):
        header = ObjectHeader(buf, address)
        self.attrs = {}
        self.dims = None
        self.dtype = None
        layout = None
        filtered = False
        for mtype, _, off, size in header.messages:
            if mtype == 0x01:
                self.dims, _ = _dataspace(buf, off)
            elif mtype == 0x03:
                self.dtype, _ = _datatype(buf, off)
            elif mtype == 0x08:
                layout = off
            elif mtype == 0x0B:
                filtered = True
            elif mtype == 0x0C:
                name, value = _attribute(buf, off)
                self.attrs[name] = value
            elif mtype == 0x15:
                fheap, = struct.unpack_from("<Q", buf, off + 2 + (2 if buf[off + 1] & 1 else 0))
                if fheap != UNDEF:
                    raise Unsupported("densely stored attributes")
        self.data = None
        if layout is None or self.dims is None or self.dtype is None:
            return
        if filtered:
            raise Unsupported("filtered (compressed) dataset")
        count = int(np.prod(self.dims)) if self.dims else 1
        version, cls = buf[layout], buf[layout + 1]
        if version == 3 and cls == 1:
            addr, nbytes = struct.unpack_from("<QQ", buf, layout + 2)
            if addr == UNDEF:
                self.data = np.zeros(self.dims, dtype=self.dtype)   # never written: fill value
            else:
                self.data = np.frombuffer(buf, dtype=self.dtype, count=count, offset=addr).reshape(self.dims).copy()
        elif version == 3 and cls == 0:
            nbytes, = struct.unpack_from("<H", buf, layout + 2)
            self.data = np.frombuffer(buf, dtype=self.dtype, count=count, offset=layout + 4).reshape(self.dims).copy()
        else:
            raise Unsupported(f"data layout version {version} class {cls}")


def _links_in_direct_block(buf, start, end):
    """Link messages packed in a fractal-heap direct block: {name: object header address}."""
    out = {}
    pos = start
    while pos + 12 < end:
        if buf[pos] == 0:                 # free space left by a removed or renamed link
            pos += 1
            continue
        if buf[pos] != 1:                 # link message version
            raise Unsupported(f"unexpected byte {buf[pos]} in a heap block at {pos}")
        flags = buf[pos + 1]
        p = pos + 2
        link_type = 0
        if flags & 0x08:
            link_type = buf[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        width = 1 << (flags & 0x3)
        n = int.from_bytes(buf[p:p + width], "little")
        p += width
        name = buf[p:p + n].decode("utf-8", "replace")
        p += n
        if link_type != 0:
            raise Unsupported("soft or external link")
        addr, = struct.unpack_from("<Q", buf, p)
        p += 8
        out[name] = addr
        pos = p
    return out


def read_file(path):
    """{dataset name: Dataset} of the root group."""
    buf = open(path, "rb").read()
    if buf[:8] != b"\x89HDF\r\n\x1a\n" or buf[8] != 0:
        raise Unsupported("not an HDF5 file with a version-0 superblock")
    if buf[13] != 8 or buf[14] != 8:
        raise Unsupported("offset/length sizes other than 8 bytes")
    root = struct.unpack_from("<Q", buf, 56 + 8)[0]
    header = ObjectHeader(buf, root)
    links = {}
    for mtype, _, off, size in header.messages:
        if mtype == 0x06:                  # compact link
            links.update(_links_in_direct_block(buf, off, off + size))
        elif mtype == 0x02:                # link info -> fractal heap
            flags = buf[off + 1]
            p = off + 2 + (8 if flags & 1 else 0)
            heap, = struct.unpack_from("<Q", buf, p)
            if heap != UNDEF:
                links.update(_dense_links(buf, heap))
    return {name: Dataset(buf, addr) for name, addr in links.items()}


def _dense_links(buf, heap):
    if buf[heap:heap + 4] != b"FRHP":
        raise Unsupported("fractal heap signature")
    # header: sig(4) ver(1) heap-id len(2) io-filter len(2) flags(1) max managed obj size(4)
    # next huge id(8) huge btree(8) free space(8) fs manager(8) managed space(8) allocated(8)
    # iterator offset(8) n managed(8) huge size(8) n huge(8) tiny size(8) n tiny(8)
    # table width(2) starting block size(8) max direct block size(8) max heap size bits(2)
    # start rows(2) root block address(8) current rows(2)
    flags = buf[heap + 9]
    p = heap + 4 + 1 + 2 + 2 + 1 + 4 + 8 * 12
    width, = struct.unpack_from("<H", buf, p)
    start_size, max_direct = struct.unpack_from("<QQ", buf, p + 2)
    max_bits, start_rows = struct.unpack_from("<HH", buf, p + 18)
    root_addr, = struct.unpack_from("<Q", buf, p + 22)
    cur_rows, = struct.unpack_from("<H", buf, p + 30)
    offset_bytes = (max_bits + 7) // 8
    checksum = 4 if flags & 0x02 else 0
    links = {}

    def direct(addr, size):
        if buf[addr:addr + 4] != b"FHDB":
            raise Unsupported("direct block signature")
        begin = addr + 4 + 1 + 8 + offset_bytes + checksum
        links.update(_links_in_direct_block(buf, begin, addr + size))

    if cur_rows == 0:
        direct(root_addr, start_size)
    else:
        if buf[root_addr:root_addr + 4] != b"FHIB":
            raise Unsupported("indirect block signature")
        q = root_addr + 4 + 1 + 8 + offset_bytes
        for row in range(cur_rows):
            size = start_size * (1 if row < 2 else 1 << (row - 1))
            for _ in range(width):
                addr, = struct.unpack_from("<Q", buf, q)
                q += 8
                if size > max_direct:
                    raise Unsupported("nested indirect blocks")
                if addr != UNDEF:
                    direct(addr, size)
    return links
