"""profiles/traffic.json from an `ncu --set full --page raw --csv` export of one step:
DRAM bytes (read + write) per launch of every kernel, keyed by the engine's kernel names
(bench.py reads it for roofline.traffic)."""
import csv
import json
import sys


def engine_name(fn):
    n = fn.split('(')[0].replace('void ', '').replace('nele::', '')
    if n.startswith('siib_jacobi2_kernel<1'):
        return 'siib_jacobi'
    if n.startswith('siib_jacobi2_kernel<4'):
        return 'siib_jacobi_cluster'
    n = n.split('<')[0]
    if n.endswith('_kernel'):
        n = n[:-7]
    return {'haspi_modcorr2': 'haspi_modcorr', 'haspi_ear_x2': 'haspi_ear', 'estoi_resample58': 'estoi_resample',
            'klt::tridiag32': 'siib_tridiag', 'bt6::backtf6': 'siib_backtf', 'bt5::backtf5': 'siib_backtf',
            'bt::backtf4': 'siib_backtf', 'qf::quadform': 'siib_quad', 'qf::quad_finish': 'siib_quad_finish'}.get(n, n)


def main(path, pairs, seconds, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    hdr, units = rows[0], rows[1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    acc = {}
    for r in rows[2:]:
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        t = acc.setdefault(engine_name(r[ik]), [0.0, 0])
        t[0] += b
        t[1] += 1
    json.dump({"pairs": pairs, "seconds": seconds, "source": path,
               "kernels": {k: v[0] / v[1] for k, v in acc.items()}}, open(out, 'w'), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4])
