"""Regenerates admm-elastic-sca_b200/csrc/glibc_exp_data.inc from this image's libm.so.6.

The table is glibc's `__exp_data` (sysdeps/ieee754/dbl-64/e_exp_data.c): invln2N, shift, negln2hiN, negln2loN, poly[4]
(C2..C5) at offsets 0x00..0x38 and tab[2*128] at offset 0xb0.  It is located by following exp()'s IFUNC resolver
(0x280d0 in Ubuntu GLIBC 2.39-0ubuntu8.5) to the FMA variant `__exp_fma` (vaddr 0x79b60), whose code loads the table base
with `lea 0x3addf(%rip)` -> 0xb4980.  If libm changes, re-derive the address with
    objdump -d /lib/x86_64-linux-gnu/libm.so.6 --start-address=<__exp_fma> | grep lea
Output: 8 + 256 IEEE-754 / integer bit patterns."""
import struct
import sys

PATH = "/lib/x86_64-linux-gnu/libm.so.6"
BASE = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0xB4980

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


head = struct.unpack("<8Q", rd(BASE, 64))
assert struct.unpack("<d", rd(BASE, 8))[0].hex() == "0x1.71547652b82fep+7", "table base moved"
tab = struct.unpack("<256Q", rd(BASE + 0xB0, 2048))
assert tab[0] == 0 and tab[1] == 0x3FF0000000000000, "tab moved"
vals = list(head) + list(tab)
for i in range(0, len(vals), 4):
    print("\t" + " ".join("0x%016xULL," % v for v in vals[i:i + 4]))
