"""Timing model of the dependency protocol of the PFRB dataflow kernel (pfnl_b200/csrc/pfrb_flow.cu): every CTA walks
its items in order, a tile runs dependency poll -> TMA loads -> MMAs -> epilogue -> publication, with the per-stage
latencies measured by tools/flow_trace.py (profiles/r2s_flow_trace_steady_state.txt).  It is a model, not a
measurement - and its one result so far was a negative one worth keeping: with the traced latencies the protocol
alone allows ~640 cycles per unit and block where the GPU needed 750-920, i.e. the kernel was NOT bound by the
dependency loop but by the throughput of its slowest role (profiles/r2z_flow_balance.txt).  TM below are
MMA issue times; the measured free-running cost of a tile is 5.3-5.4 K (3x3 roles) and 7.0 K (conv10).
    python tools/flow_sim.py [clips] [tiles_y] [tiles_x] [--split a,b,c,d] [--look n] [--jitter f]"""
import argparse

ROLES = ("conv1", "conv10", "conv2b", "conv2f")
# cycles: MMA issue time per tile, accumulator ready -> stores issued, stores issued -> counter visible,
# producer go -> data in shared memory, dependency seen -> producer go
TM = {"conv1": 4200, "conv10": 5700, "conv2b": 4200, "conv2f": 4500}
E1 = {"conv1": 2300, "conv10": 1500, "conv2b": 1600, "conv2f": 2700}
PUB = {"conv1": 1800, "conv10": 2300, "conv2b": 1900, "conv2f": 2600}
LLOAD = {"conv1": 1100, "conv10": 1800, "conv2b": 1400, "conv2f": 2100}
LDEP = 950
WSWAP = 3000  # weights of the next block: drain + load
FR = 7


def simulate(n_img, ty, tx, split, nblk=20, look=1, jitter=0.0, seed=1):
    import random
    rng = random.Random(seed)
    U = n_img * ty * tx
    short_y = ty < tx

    def coords(u):
        if short_y:
            y = u % ty
            r = u // ty
            return r // tx, y, r % tx
        x = u % tx
        r = u // tx
        return r // ty, r % ty, x

    def unit(n, y, x):
        return (n * tx + x) * ty + y if short_y else (n * ty + y) * tx + x

    nbrs = []
    for u in range(U):
        n, y, x = coords(u)
        nbrs.append([unit(n, y + dy, x + dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
                     if 0 <= y + dy < ty and 0 <= x + dx < tx])
    n_role = dict(zip(ROLES, split))
    pub = {r: {} for r in ROLES}  # (b, u[, t]) -> publication time

    def deps(role, b, u, t):
        if role == "conv10":
            return [("conv1", (b, u, f)) for f in range(FR)]
        if role == "conv2f":
            return [("conv2b", (b, u))]
        if role == "conv2b":
            return [("conv10", (b, v)) for v in nbrs[u]]
        return [("conv2f", (b - 1, v, t)) for v in nbrs[u]] if b > 0 else []

    ctas = []
    for role in ROLES:
        per_frame = role in ("conv1", "conv2f")
        n_items = U * FR if per_frame else U
        for rank in range(n_role[role]):
            items = []
            for b in range(nblk):
                for it in range(rank, n_items, n_role[role]):
                    items.append((b, it // FR, it % FR) if per_frame else (b, it, 0))
            ctas.append({"role": role, "items": items, "pos": 0, "go": [], "mma_s": [], "mma_e": [], "epi_e": [],
                         "wait": 0.0, "per_frame": per_frame})
    done = 0
    total = sum(len(c["items"]) for c in ctas)
    while done < total:
        progressed = False
        for c in ctas:
            role = c["role"]
            while c["pos"] < len(c["items"]):
                b, u, t = c["items"][c["pos"]]
                ds = deps(role, b, u, t)
                try:
                    dep_t = max((pub[r][k] for r, k in ds), default=0.0)
                except KeyError:
                    break
                i = c["pos"]
                # the producer may run `look` tiles ahead of the MMA warp (ring depth)
                go = dep_t + LDEP if ds else 0.0
                if i >= look:
                    go = max(go, c["mma_s"][i - look])
                ready = go + LLOAD[role]
                start = ready
                if i >= 1:
                    gap = 400
                    if c["items"][i - 1][0] != b:
                        gap = WSWAP
                    start = max(start, c["mma_e"][i - 1] + gap)
                    if start > c["mma_e"][i - 1] + gap:
                        c["wait"] += start - (c["mma_e"][i - 1] + gap)
                if i >= 2:
                    start = max(start, c["epi_e"][i - 2])
                end = start + TM[role] * (1.0 + jitter * rng.random())
                epi_s = end + 170
                if i >= 1:
                    epi_s = max(epi_s, c["epi_e"][i - 1])
                epi_e = epi_s + E1[role]
                c["go"].append(go)
                c["mma_s"].append(start)
                c["mma_e"].append(end)
                c["epi_e"].append(epi_e)
                key = (b, u, t) if c["per_frame"] else (b, u)
                pub[role][key] = epi_e + PUB[role]
                c["pos"] += 1
                done += 1
                progressed = True
        if not progressed:
            raise RuntimeError("deadlock in the model")
    span = max(max(v.values()) for v in pub.values())
    res = {"units": U, "cycles": span, "cycles_per_unit_block": span / (U * nblk), "ms_at_1.92GHz": span / 1.92e6}
    for role in ROLES:
        cs = [c for c in ctas if c["role"] == role]
        ends = [c["epi_e"][-1] for c in cs]
        res[role + "_end_spread"] = round(max(ends) - min(ends))
        busy = sum(len(c["items"]) * TM[role] for c in cs) / (len(cs) * span)
        res[role + "_busy"] = round(busy, 3)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("clips", type=int, nargs="?", default=16)
    ap.add_argument("tiles_y", type=int, nargs="?", default=2)
    ap.add_argument("tiles_x", type=int, nargs="?", default=4)
    ap.add_argument("--split", default="64,10,10,64")
    ap.add_argument("--look", type=int, default=1)
    ap.add_argument("--jitter", type=float, default=0.0)
    a = ap.parse_args()
    print(simulate(a.clips, a.tiles_y, a.tiles_x, [int(v) for v in a.split.split(",")], look=a.look, jitter=a.jitter))
