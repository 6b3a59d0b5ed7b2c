"""tests/golden/vlp16_golden.npz -> the flat binary case file read by ref_dump (and by tests/c_abi/adapter_driver.cc).
    python oracle/ref_harness/export_case.py tests/golden/vlp16_golden.npz /tmp/case.bin"""
import struct
import sys

import numpy as np


def write_case(path, arrays, ring_lc, ring_ls, init_map, init_odo):
    with open(path, "wb") as f:
        f.write(struct.pack("8i", *[a.shape[0] for a in arrays]))
        for a in arrays:
            f.write(np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4).tobytes())
        f.write(np.asarray(ring_lc, dtype=np.float32).tobytes())
        f.write(np.asarray(ring_ls, dtype=np.float32).tobytes())
        f.write(np.asarray(init_map, dtype=np.float64).tobytes())
        f.write(np.asarray(init_odo, dtype=np.float64).tobytes())


if __name__ == "__main__":
    g = np.load(sys.argv[1])
    write_case(sys.argv[2],
               [g["map_corner"], g["map_surf"], g["scan_corner"], g["scan_surf"],
                g["odo_last_corner"], g["odo_last_surf"], g["odo_curr_sharp"], g["odo_curr_flat"]],
               g["odo_last_corner_ring"], g["odo_last_surf_ring"], g["init"], g["odo_init"])
    print("wrote", sys.argv[2])
