"""Print the kernels of the LAST step of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_list.py gpurun_out/launches.csv [first-kernel-substring]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    first = sys.argv[2] if len(sys.argv) > 2 else 'tc_pack_batched'
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    names = [(r[ki][:84], float(r[vi].replace(',', '')) / 1000) for r in rows[start + 2:] if len(r) > vi]
    idx = [i for i, (n, _) in enumerate(names) if first in n]
    last = names[idx[-1]:] if idx else names
    tot = 0.0
    for n, t in last:
        print('%8.1f us  %s' % (t, n))
        tot += t
    print('%d launches, %.1f us summed (cold-cache, serialised)' % (len(last), tot))


if __name__ == '__main__':
    main()
