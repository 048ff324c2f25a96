"""Bus bandwidth vs message size, ours against NCCL, as dependency-free SVG (no matplotlib in the image).
Counterpart of the reference's test/host/Coyote/run_scripts/plot.py (throughput curves with an MPI overlay).

  python bench/plot.py profiles/sweep_8gpu_nvls.csv [more.csv ...]      # writes <csv stem>.svg next to each CSV
"""
import csv
import math
import os
import sys

COLORS = {"accl": "#1f77b4", "nccl": "#d62728"}
W, H, ML, MR, MT, MB = 420, 300, 56, 12, 34, 44


def human(n):
    for u in ("B", "K", "M", "G"):
        if n < 1024:
            return f"{n:g}{u}"
        n /= 1024
    return f"{n:g}T"


def panel(op, rows, x0, y0, world, dtype):
    xs = [math.log2(int(r["bytes"])) for r in rows]
    ys_a = [float(r["accl_busbw"]) for r in rows]
    ys_n = [float(r["nccl_busbw"]) if r.get("nccl_busbw") else None for r in rows]
    ymax = max([y for y in ys_a + [v for v in ys_n if v] if y] + [1.0])
    lo, hi = math.log10(max(min(y for y in ys_a if y > 0), 1e-2)), math.log10(ymax * 1.3)
    xmin, xmax = min(xs), max(xs)

    def px(x):
        return x0 + ML + (x - xmin) / max(xmax - xmin, 1e-9) * (W - ML - MR)

    def py(y):
        return y0 + MT + (1 - (math.log10(max(y, 1e-2)) - lo) / max(hi - lo, 1e-9)) * (H - MT - MB)

    out = [f'<rect x="{x0 + ML}" y="{y0 + MT}" width="{W - ML - MR}" height="{H - MT - MB}" fill="white" stroke="#888"/>',
           f'<text x="{x0 + W / 2}" y="{y0 + 20}" text-anchor="middle" font-size="13" font-weight="bold">{op} — {world} x B200, {dtype}</text>']
    for e in range(math.floor(lo), math.ceil(hi) + 1):  # y grid: decades
        y = 10.0 ** e
        if lo <= e <= hi:
            out.append(f'<line x1="{x0 + ML}" x2="{x0 + W - MR}" y1="{py(y):.1f}" y2="{py(y):.1f}" stroke="#ddd"/>')
            out.append(f'<text x="{x0 + ML - 4}" y="{py(y) + 4:.1f}" text-anchor="end" font-size="10">{y:g}</text>')
    for x in range(int(xmin), int(xmax) + 1, 2):       # x ticks: every 4x
        out.append(f'<line x1="{px(x):.1f}" x2="{px(x):.1f}" y1="{y0 + MT}" y2="{y0 + H - MB}" stroke="#eee"/>')
        out.append(f'<text x="{px(x):.1f}" y="{y0 + H - MB + 13}" text-anchor="middle" font-size="10">{human(2 ** x)}</text>')
    out.append(f'<text x="{x0 + 12}" y="{y0 + H / 2}" font-size="10" transform="rotate(-90 {x0 + 12} {y0 + H / 2})" text-anchor="middle">bus GB/s</text>')
    for name, ys in (("nccl", ys_n), ("accl", ys_a)):
        pts = [(px(x), py(y)) for x, y in zip(xs, ys) if y]
        if not pts:
            continue
        out.append(f'<polyline fill="none" stroke="{COLORS[name]}" stroke-width="2" points="' +
                   " ".join(f"{a:.1f},{b:.1f}" for a, b in pts) + '"/>')
        out += [f'<circle cx="{a:.1f}" cy="{b:.1f}" r="2.5" fill="{COLORS[name]}"/>' for a, b in pts]
    lx, ly = x0 + ML + 8, y0 + MT + 14
    out.append(f'<line x1="{lx}" x2="{lx + 18}" y1="{ly - 4}" y2="{ly - 4}" stroke="{COLORS["accl"]}" stroke-width="2"/><text x="{lx + 22}" y="{ly}" font-size="10">accl_b200</text>')
    out.append(f'<line x1="{lx}" x2="{lx + 18}" y1="{ly + 10}" y2="{ly + 10}" stroke="{COLORS["nccl"]}" stroke-width="2"/><text x="{lx + 22}" y="{ly + 14}" font-size="10">NCCL</text>')
    return out


def main():
    for path in sys.argv[1:]:
        rows = list(csv.DictReader(open(path)))
        if not rows:
            continue
        ops = []
        for r in rows:
            if r["op"] not in ops:
                ops.append(r["op"])
        cols = min(3, len(ops))
        nrow = (len(ops) + cols - 1) // cols
        body = []
        for i, op in enumerate(ops):
            sel = sorted((r for r in rows if r["op"] == op), key=lambda r: int(r["bytes"]))
            body += panel(op, sel, (i % cols) * W, (i // cols) * H, rows[0]["world"], rows[0]["dtype"])
        svg = (f'<svg xmlns="http://www.w3.org/2000/svg" width="{cols * W}" height="{nrow * H}" font-family="sans-serif">'
               f'<rect width="100%" height="100%" fill="#fafafa"/>' + "".join(body) + "</svg>")
        out = os.path.splitext(path)[0] + ".svg"
        open(out, "w").write(svg)
        print(out)


if __name__ == "__main__":
    main()
