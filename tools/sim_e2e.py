"""Back-of-the-envelope model of ozimmu_gemm_host's block pipeline (no GPU needed): python tools/sim_e2e.py

Blocks of op(A) and op(B) arrive alternately over PCIe; the rectangle of C an arrival completes is handed to the GPU,
which works through the rectangles in order at a fixed rate (`fluid`: perfectly divisible work, i.e. what a shared
tile queue would give; `rounds`: one launch per rectangle, each rounded up to whole rounds of 74 tiles, launches
strictly one after the other -- the worst case the stream rotation is there to avoid).  Prints, per schedule, when the
last byte lands, when the last product ends and when the last byte of C is back on the host.  The continuous limit of
the alternating schedule (compute starved until t* = T^2 / 2 Tc, saturated afterwards) is printed for comparison.
DESIGN.md section 4 quotes these numbers next to the measured ones (profiles/r1_e2e_block_sweep.txt)."""
import math

N = K = 8192
H2D = 55.0e9              # B/s, measured (profiles/r1_e2e_block_sweep.txt: 512 MiB in 9.75 ms)
D2H = 52.0e9
T_COMPUTE = 17.1e-3       # s, one 8192^3 fp64_int8_9 product, device resident
PAIRS = 74
T_ROUND = T_COMPUTE / (1024 / PAIRS)


def blocks(extent, edge, tail=()):
    out, at = [], 0
    body = extent - sum(tail)
    while body - at > edge:
        out.append(edge)
        at += edge
    out.append(body - at)
    return out + list(tail)


def simulate(ha, wb, mode):
    t, order, ia, ib = 0.0, [], 0, 0
    while ia < len(ha) or ib < len(wb):
        if ib < len(wb) and (ib <= ia or ia >= len(ha)):
            t += wb[ib] * K * 8 / H2D
            order.append((t, "B", ib))
            ib += 1
        else:
            t += ha[ia] * K * 8 / H2D
            order.append((t, "A", ia))
            ia += 1
    gpu = out = 0.0
    na = nb = 0
    for arrived, kind, idx in order:
        if kind == "A":
            rows, cols = ha[idx], sum(wb[:nb])
            na += 1
        else:
            rows, cols = sum(ha[:na]), wb[idx]
            nb += 1
        gpu = max(gpu, arrived) + 0.05e-3                      # the block's split
        if rows and cols:
            tiles = math.ceil(rows / 256) * math.ceil(cols / 256)
            work = tiles / PAIRS if mode == "fluid" else math.ceil(tiles / PAIRS)
            gpu += work * T_ROUND
            out = max(out, gpu) + rows * cols * 8 / D2H
    return t * 1e3, gpu * 1e3, out * 1e3


if __name__ == "__main__":
    T = 2 * N * K * 8 / H2D
    t_star = T * T / (2 * T_COMPUTE)
    done = (t_star / T) ** 2 + (T - t_star) / T_COMPUTE
    print(f"continuous limit: last byte at {T*1e3:.1f} ms, compute saturated from {t_star*1e3:.1f} ms, "
          f"{(1-done)*100:.0f} % of the work left at the last byte -> {(T + (1-done)*T_COMPUTE)*1e3:.1f} ms + last copy-out")
    for name, edge, tail in [("1024", 1024, ()), ("768", 768, ()), ("512", 512, ()), ("2048", 2048, ()),
                             ("1024 tapered", 1024, (512, 256, 256)), ("768 tapered", 768, (256, 256))]:
        b = blocks(N, edge, tail)
        for mode in ("fluid", "rounds"):
            arrive, comp, out = simulate(b, b, mode)
            print(f"blocks {name:13s} {mode:6s}: last byte {arrive:5.1f} ms, last product {comp:5.1f} ms, C on the host {out:5.1f} ms")
