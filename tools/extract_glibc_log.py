"""Regenerates admm-elastic-sca_b200/csrc/glibc_log_data.h from this image's libm.so.6.

The table is glibc's `__log_data` (ln2hi, ln2lo, A[5], B[11], tab[128]{invc,logc}); it is located by following
log()'s IFUNC resolver to the FMA variant (vaddr 0x79d50 in Ubuntu GLIBC 2.39-0ubuntu8.5) whose code loads the
table base with `lea 0x3b4a0(%rip)` -> 0xb5240.  If libm changes, re-derive the address with
    objdump -d /lib/x86_64-linux-gnu/libm.so.6 --start-address=<__log_fma> | grep lea
"""
import struct
import sys

PATH = "/lib/x86_64-linux-gnu/libm.so.6"
BASE = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0xB5240
N = 2 + 5 + 11 + 256

data = open(PATH, "rb").read()
e_phoff = struct.unpack_from("<Q", data, 0x20)[0]
e_phentsize, e_phnum = struct.unpack_from("<HH", data, 0x36)
segs = []
for i in range(e_phnum):
    p_type, _, p_offset, p_vaddr, _, p_filesz, _, _ = struct.unpack_from("<IIQQQQQQ", data, e_phoff + i * e_phentsize)
    if p_type == 1:
        segs.append((p_vaddr, p_offset, p_filesz))


def rd(vaddr, n):
    for va, off, sz in segs:
        if va <= vaddr < va + sz:
            return data[off + vaddr - va: off + vaddr - va + n]
    raise KeyError(hex(vaddr))


vals = struct.unpack("<%dQ" % N, rd(BASE, 8 * N))
fl = struct.unpack("<%dd" % N, rd(BASE, 8 * N))
assert fl[0].hex() == "0x1.62e42fefa3800p-1", "table base moved: " + fl[0].hex()
for v in vals:
    print("0x%016x" % v)
