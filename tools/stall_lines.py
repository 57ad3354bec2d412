"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line: the lines with the most warp-stall
samples per kernel, with their dominant stall reasons.  Usage: python tools/stall_lines.py prof_src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == 'Function Name':
        name = r[1]
        hdr = rows[i + 1]
        j = i + 2
        lines = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ('File Path', 'Function Name')):
            if len(rows[j]) == len(hdr):
                lines.append(rows[j])
            j += 1
        # CUDA-source rows have a line number and '-' address
        ci = {h: k for k, h in enumerate(hdr)}
        samp = ci['# Samples']
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith('stall_') and '(Not Issued)' not in h]
        src_rows = [l for l in lines if l[2] == '-' and l[0].isdigit()]
        total = sum(int(l[samp] or 0) for l in src_rows) or 1
        inst = sum(int(l[ci['Instructions Executed']] or 0) for l in src_rows)
        print('== %s  (%d stall samples, %d warp instructions)' % (name[:90], total, inst))
        for l in sorted(src_rows, key=lambda l: -int(l[samp] or 0))[:top]:
            st = sorted(((int(l[k] or 0), hdr[k][6:]) for k in stall_cols), reverse=True)[:3]
            print('  %5.1f%%  L%-4s %-100s  [%s]' % (100.0 * int(l[samp] or 0) / total, l[0], l[1].strip()[:100],
                                                  ', '.join('%s %d' % (n, c) for c, n in st if c)))
        i = j
    else:
        i += 1
