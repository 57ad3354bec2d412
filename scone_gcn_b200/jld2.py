"""Minimal pure-Python reader for the JLD2 / HDF5 subset used by ocean_drifters_data/dataBuoys.jld2 (h5py is not available).

Supports exactly what that file needs (ocean_drifters_data/buoy_data.py:12-36 reads `elist`, `tlist`, `TrajectoriesNodes`):
HDF5 superblock v2 behind the 512-byte JLD2 text header, v2 object headers (+ continuation blocks), link messages in the
root group header, simple dataspaces, fixed-point / reference datatypes, compact and contiguous layouts, and 8-byte object
references (dereferenced like `f[ref][()]`).  Anything else raises NotImplementedError.
"""
import struct

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'


class JLD2File:
    def __init__(self, path):
        with open(path, 'rb') as f:
            self.buf = f.read()
        self.base = self.buf.find(_SIG)
        if self.base < 0:
            raise ValueError('no HDF5 superblock found')
        sb = self.base
        version = self.buf[sb + 8]
        if version not in (2, 3):
            raise NotImplementedError('HDF5 superblock version %d' % version)
        if self.buf[sb + 9] != 8 or self.buf[sb + 10] != 8:
            raise NotImplementedError('offset/length size other than 8')
        self.base_addr, _ext, _eof, root = struct.unpack_from('<QQQQ', self.buf, sb + 12)
        self.links = {}
        for mtype, data in self._messages(root):
            if mtype == 6:
                name, addr = self._link(data)
                self.links[name] = addr

    # ---- low level ---------------------------------------------------------------------------------------------
    def _abs(self, addr):
        return self.base_addr + addr

    def _messages(self, addr):
        """Yield (type, bytes) for every header message of the v2 object header at relative address `addr`."""
        p = self._abs(addr)
        if self.buf[p:p + 4] != b'OHDR' or self.buf[p + 4] != 2:
            raise NotImplementedError('only version-2 object headers are supported')
        flags = self.buf[p + 5]
        p += 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        chunk = int.from_bytes(self.buf[p:p + nsz], 'little')
        p += nsz
        blocks = [(p, p + chunk)]
        tracked = bool(flags & 0x04)
        while blocks:
            lo, hi = blocks.pop(0)
            while lo + 4 <= hi:
                mtype = self.buf[lo]
                size = struct.unpack_from('<H', self.buf, lo + 1)[0]
                lo += 4 + (2 if tracked else 0)
                data = self.buf[lo:lo + size]
                lo += size
                if mtype == 0x10:                                   # continuation: OCHK block, trailing checksum
                    off, length = struct.unpack_from('<QQ', data, 0)
                    q = self._abs(off)
                    if self.buf[q:q + 4] != b'OCHK':
                        raise ValueError('bad continuation block')
                    blocks.append((q + 4, q + length - 4))
                elif mtype != 0:
                    yield mtype, data

    @staticmethod
    def _link(d):
        if d[0] != 1:
            raise NotImplementedError('link message version %d' % d[0])
        flags = d[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[p]
            p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nsz = 1 << (flags & 3)
        nlen = int.from_bytes(d[p:p + nsz], 'little')
        p += nsz
        name = d[p:p + nlen].decode('utf-8')
        p += nlen
        if ltype != 0:
            raise NotImplementedError('soft / external links')
        return name, struct.unpack_from('<Q', d, p)[0]

    def _dataset(self, addr):
        """Returns a numpy array: int64 data, or uint64 object references (dtype '<u8' with .is_ref attribute)."""
        dims, dt_class, dt_size, raw = (), None, None, None
        for mtype, d in self._messages(addr):
            if mtype == 1:                                          # dataspace
                ver, rank, fl = d[0], d[1], d[2]
                p = 8 if ver == 1 else 4
                dims = struct.unpack_from('<%dQ' % rank, d, p) if rank else ()
            elif mtype == 3:                                        # datatype
                dt_class, dt_size = d[0] & 0x0f, struct.unpack_from('<I', d, 4)[0]
            elif mtype == 8:                                        # layout
                ver, cls = d[0], d[1]
                if ver not in (3, 4):
                    raise NotImplementedError('layout version %d' % ver)
                if cls == 0:
                    size = struct.unpack_from('<H', d, 2)[0]
                    raw = d[4:4 + size]
                elif cls == 1:
                    a, size = struct.unpack_from('<QQ', d, 2)
                    raw = self.buf[self._abs(a):self._abs(a) + size] if a != 0xffffffffffffffff else b''
                else:
                    raise NotImplementedError('chunked layout')
        if raw is None or dt_class is None:
            raise ValueError('object at %d is not a simple dataset' % addr)
        n = int(np.prod(dims)) if dims else 1
        if dt_class == 0 and dt_size == 8:                          # fixed-point Int64
            arr = np.frombuffer(raw, dtype='<i8', count=n)
            return arr.reshape(dims) if dims else arr[0], False             # same C-order view h5py gives (f['elist'][:])
        if dt_class == 7:                                           # object references
            arr = np.frombuffer(raw, dtype='<u8', count=n)
            return arr.reshape(dims) if dims else arr[0], True
        raise NotImplementedError('datatype class %d size %d' % (dt_class, dt_size))

    # ---- public ----------------------------------------------------------------------------------------------------
    def keys(self):
        return list(self.links)

    def read(self, name):
        """Dataset `name` of the root group; arrays of references are dereferenced recursively (lists of values)."""
        return self._resolve(self.links[name])

    def _resolve(self, addr):
        val, is_ref = self._dataset(addr)
        if not is_ref:
            return val
        if np.ndim(val) == 0:
            return self._resolve(int(val))
        return [self._resolve(int(r)) for r in np.asarray(val).ravel()]
