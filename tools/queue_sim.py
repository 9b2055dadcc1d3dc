#!/usr/bin/env python
"""Scheduling simulator for the per-warp item queues of the warp-queue render kernel (K3, render_kernels.cu).

Runs the kernel's queue discipline - refill, root test, `drain` with its batch priorities and overflow guard, shading - for
the rays of a few 8x4 tiles on the CPU, on the oracle's own tree, and counts node / leaf batches and how full they are, for
the two pop orders of the node queue:

  lifo    every node batch takes the NEWEST 32 items (round 1 / first half of round 2)
  deque   batches take the OLDEST items while at most T are queued, the newest above that

It is a development tool (needs no GPU): it predicted the lane use ncu then measured (rgbbox, 64 rays per warp: 0.87 with
`lifo`, 0.98 with `deque`, T = 128; ncu: 27.7 -> 30.2 of 32 lanes in the box tests) and the share of node batches without a
leaf child (2 in 3).  Arithmetic is float64: only the SCHEDULING statistics are meaningful, pixels are not produced.
The oracle (test infrastructure) is used for the scene, the LBVH and the camera.

  python tools/queue_sim.py rgbbox 64 64 [tiles]     scene, spp, rays per warp (32 * K), number of random tiles
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O  # noqa: E402

NODE_BATCH_INSTR, LEAF_BATCH_INSTR = 124, 95   # SASS instructions of the full batch bodies (cost model of the summary line)


def setup(name, h, w):
    sc = O.Scene.named(name)
    d = sc.prepare(h, w).dump()
    spheres, _ = sc.arrays()
    return dict(left=d["left"], right=d["right"], boxes=d["boxes"].astype(np.float64), rootbox=d["boxes"][0].astype(np.float64),
                sph=spheres[d["perm"]].astype(np.float64), cam=d["cam"].astype(np.float64), h=h, w=w)   # spheres in Morton order


def box_hit(b, o, inv):
    tmin, tmax = 0.0, 1e9
    for a in range(3):
        t0, t1 = (b[a] - o[a]) * inv[a], (b[3 + a] - o[a]) * inv[a]
        if inv[a] < 0:
            t0, t1 = t1, t0
        if t0 == t0:
            tmin = max(tmin, t0)
        if t1 == t1:
            tmax = min(tmax, t1)
    return tmax > tmin


def sphere_t(sp, o, d, tmin, tmax):
    oc = o - sp[:3]
    a, b, c = d @ d, oc @ d, oc @ oc - sp[3] * sp[3]
    disc = b * b - a * c
    if not disc > 0:
        return -1.0
    sq = math.sqrt(disc)
    for root in ((-b - sq) / a, (-b + sq) / a):
        if tmin < root < tmax:
            return root
    return -1.0


def primary(S, pi, pj, smp):
    cam = S["cam"]   # origin, lower-left corner, horizontal, vertical
    ox = (smp * 0.7548776662) % 1.0 if smp else 0.0
    oy = (smp * 0.5698402909) % 1.0 if smp else 0.0
    u, v = (pi + ox) / S["w"], (S["h"] - pj + oy) / S["h"]
    return cam[0:3].copy(), cam[3:6] + u * cam[6:9] + v * cam[9:12] - cam[0:3]


class Stats:
    def __init__(self):
        self.rounds = self.nb_full = self.nb_part = self.lb_full = self.lb_part = 0
        self.n_items = self.l_items = self.part_fill = self.no_leaf_child = 0

    def cost(self):
        return (self.nb_full + self.nb_part) * NODE_BATCH_INSTR + (self.lb_full + self.lb_part) * LEAF_BATCH_INSTR


def run_warp(S, pixels, spp, R, policy, T, ncap, st):
    """One warp's life: `pixels` x spp samples, dispensed in order to R slots; rounds as in the kernel."""
    left, right, boxes, sph = S["left"], S["right"], S["boxes"], S["sph"]
    work = [(p, s) for p in pixels for s in range(spp)]
    wpos, slots, nroom = 0, [None] * R, ncap - 96
    while True:
        go, started = [], set()
        for _ in range(4):                                    # refill passes: sky rays never occupy a traversal round
            for i in range(R):
                if slots[i] is None and wpos < len(work):
                    (pi, pj), s = work[wpos]
                    wpos += 1
                    o, d = primary(S, pi, pj, s)
                    slots[i] = [o, d, 0]
            for i in range(R):
                if i in started or slots[i] is None:
                    continue
                o, d, depth = slots[i]
                inv = 1.0 / np.where(d == 0, 1e-300, d)
                if box_hit(S["rootbox"], o, inv):
                    started.add(i)
                    go.append(i)
                    slots[i] = [o, d, depth, inv]
                else:
                    slots[i] = None                           # root miss: sky
            if wpos >= len(work) or all(s is not None for s in slots):
                break
        if not go:
            if wpos >= len(work) and all(s is None for s in slots):
                return
            continue
        st.rounds += 1
        best = {i: (1e30, -1) for i in go}
        nst, lst = [(i, 0) for i in go], []                   # list end = top of the queue

        def node_batch(n, from_bottom=False):
            nonlocal nst
            if from_bottom:
                items, nst = nst[:n], nst[n:]
            else:
                items, nst = nst[len(nst) - n:][::-1], nst[:len(nst) - n]
            pushes_n, pushes_l = [], []
            for i, cur in items:
                o, _, _, inv = slots[i]
                for ch in (left[cur], right[cur]):
                    if ch < 0:
                        pushes_l.append((i, ~ch))             # a leaf child has no box in the reference: always visited
                    elif box_hit(boxes[ch], o, inv):
                        pushes_n.append((i, ch))
            nst.extend(pushes_n[::-1])
            lst.extend(pushes_l)
            st.n_items += n
            st.no_leaf_child += not pushes_l
            if n == 32:
                st.nb_full += 1
            else:
                st.nb_part += 1
                st.part_fill += n

        def leaf_batch(n):
            nonlocal lst
            items, lst = lst[len(lst) - n:], lst[:len(lst) - n]
            for i, li in items:
                o, d, _, _ = slots[i]
                t = sphere_t(sph[li], o, d, 0.1, 1e9)
                if t >= 0 and (t, li) < best[i]:
                    best[i] = (t, li)
            st.l_items += n
            if n == 32:
                st.lb_full += 1
            else:
                st.lb_part += 1

        while True:                                           # drain
            while 32 <= len(nst) <= nroom and len(lst) < 32:
                node_batch(32, from_bottom=policy == "deque" and len(nst) <= T)
            if len(lst) >= 32:
                leaf_batch(32)
            elif nst:
                node_batch(1 if len(nst) > nroom else min(len(nst), 32))
            elif lst:
                leaf_batch(len(lst))
            else:
                break
        for i in go:                                          # shade
            o, d, depth, _ = slots[i]
            t, li = best[i]
            if li < 0:
                slots[i] = None
                continue
            sp = sph[li]
            p = o + sphere_t(sp, o, d, 0.0, t + 1) * d
            n = (p - sp[:3]) / sp[3]
            v = d / math.sqrt(d @ d)
            refl = v - 2 * (v @ n) * n
            slots[i] = [p, refl, depth + 1] if refl @ n > 0 and depth + 1 < 50 else None


def main():
    name, spp, R = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    ntiles = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    h = w = 1000
    S = setup(name, h, w)
    tiles_x, ntile = w // 8, (h // 4) * (w // 8)
    ncap = 256 if R > 32 else 512
    for policy, T in (("lifo", 0), ("deque", ncap // 4), ("deque", ncap // 2)):
        st, rng = Stats(), np.random.default_rng(1)
        for _ in range(ntiles):
            t = int(rng.integers(ntile * 45 // 100 if name == "irreg" else 0, ntile))   # irreg: skip the sky rows
            ty, tx = divmod(t, tiles_x)
            run_warp(S, [(tx * 8 + (k % 8), ty * 4 + k // 8) for k in range(32)], spp, R, policy, T, ncap, st)
        nb = st.nb_full + st.nb_part
        print(f"{name} spp {spp} R {R} {policy:5s} T {T:3d}: rounds {st.rounds}, node batches {nb} (full {st.nb_full}, partial "
              f"{st.nb_part}, mean fill {st.part_fill / max(st.nb_part, 1):.1f}), lane use {st.n_items / 32 / max(nb, 1):.3f}, "
              f"without a leaf child {st.no_leaf_child}; leaf batches {st.lb_full}+{st.lb_part}; cost {st.cost() / 1e3:.0f}k "
              f"(all batches full: {(st.n_items * NODE_BATCH_INSTR + st.l_items * LEAF_BATCH_INSTR) / 32e3:.0f}k)", flush=True)


if __name__ == "__main__":
    main()
