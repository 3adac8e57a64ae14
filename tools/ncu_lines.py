"""Per-source-line warp-stall samples of one kernel from an ncu report captured with
--set full --import-source on (built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source',
                          'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file = ''
    hdr = None
    lines = {}            # (file, line) -> [source, samples, not_issued, instr, {stall: n}]
    key = None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            si = hdr.index('# Samples')
            ni = hdr.index('Warp Stall Sampling (Not-issued Samples)')
            ii = hdr.index('Instructions Executed')
            st0 = hdr.index('stall_barrier')
            st1 = hdr.index('stall_barrier (Not Issued)')
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != '':
            key = (cur_file, int(r[0]))
            lines.setdefault(key, [r[1], 0, 0, 0, {}])
            continue                      # per-line totals repeat the SASS rows below
        e = lines[key]
        num = lambda t: int(t) if t.strip().isdigit() else 0
        e[1] += num(r[si])
        e[2] += num(r[ni])
        e[3] += num(r[ii])
        for j in range(st0, st1):
            v = num(r[j])
            if v:
                e[4][hdr[j]] = e[4].get(hdr[j], 0) + v
    tot = sum(e[1] for e in lines.values())
    print(f'total samples {tot}')
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ' '.join(f'{k[6:]}:{v}' for k, v in sorted(e[4].items(), key=lambda kv: -kv[1])[:3])
        print(f'{100.0 * e[1] / max(tot, 1):5.1f}% {e[1]:7d} inst {e[3]:9d}  {f}:{ln:<4d} {e[0].strip()[:70]:70s} | {st}')


if __name__ == '__main__':
    main()
