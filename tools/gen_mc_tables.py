#!/usr/bin/env python
"""Generates the marching-cubes case table (surf_b200/csrc/mc_tables.cuh and oracle/mc_tables.py) from first
principles — no third-party table is copied.

Cube corner c = cx + 2 cy + 4 cz.  Edge e = axis * 4 + j joins corners that differ along `axis`; j enumerates the
other two coordinates (lower-axis bit first).  A case is the 8-bit set of INSIDE corners.  For every case:
  1. on each of the 6 faces connect the crossing edges pairwise; a face with 4 crossing edges (two diagonal inside
     corners) is resolved by cutting off each INSIDE corner separately — the rule depends only on the face's own 4
     corner states, so the two cubes sharing a face always agree and the surface is watertight;
  2. the face segments form closed loops over the crossing edges (each crossing edge lies on exactly two faces);
  3. every loop is fan-triangulated and oriented so that the normal points from inside to outside.
"""
import itertools
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corner_xyz(c):
    return np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1], dtype=np.float64)


def build_edges():
    edges = []
    for axis in range(3):
        others = [a for a in range(3) if a != axis]
        for j in range(4):
            base = [0, 0, 0]
            base[others[0]] = j & 1
            base[others[1]] = (j >> 1) & 1
            a = base[0] + 2 * base[1] + 4 * base[2]
            b = a + (1 << axis)
            edges.append((a, b))
    return edges


EDGES = build_edges()
EDGE_OF = {frozenset(e): i for i, e in enumerate(EDGES)}


def faces():
    """6 faces as cyclic corner quadruples."""
    out = []
    for axis in range(3):
        o0, o1 = [a for a in range(3) if a != axis]
        for side in (0, 1):
            quad = []
            for (u, v) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[axis] = side
                p[o0] = u
                p[o1] = v
                quad.append(p[0] + 2 * p[1] + 4 * p[2])
            out.append(quad)
    return out


FACES = faces()


def face_segments(case, quad):
    inside = [(case >> c) & 1 for c in quad]
    fe = [EDGE_OF[frozenset((quad[i], quad[(i + 1) % 4]))] for i in range(4)]      # edge i: corner i -> i+1
    cross = [i for i in range(4) if inside[i] != inside[(i + 1) % 4]]
    if len(cross) == 0:
        return []
    if len(cross) == 2:
        return [(fe[cross[0]], fe[cross[1]])]
    # 4 crossings: corners alternate; cut off each inside corner: corner i sits between edges i-1 and i
    segs = []
    for i in range(4):
        if inside[i]:
            segs.append((fe[(i - 1) % 4], fe[i]))
    return segs


def edge_mid(e):
    a, b = EDGES[e]
    return 0.5 * (corner_xyz(a) + corner_xyz(b))


def trilinear_grad(case, p):
    """Gradient at p of the trilinear interpolant of the +1 (inside) / -1 (outside) corner field."""
    g = np.zeros(3)
    for c in range(8):
        v = 1.0 if (case >> c) & 1 else -1.0
        q = corner_xyz(c)
        w = [(p[a] if q[a] else 1 - p[a]) for a in range(3)]
        for a in range(3):
            d = (1.0 if q[a] else -1.0)
            g[a] += v * d * np.prod([w[b] for b in range(3) if b != a])
    return g


def triangulate(case):
    adj = {}
    for quad in FACES:
        for (e0, e1) in face_segments(case, quad):
            adj.setdefault(e0, []).append(e1)
            adj.setdefault(e1, []).append(e0)
    for e, nb in adj.items():
        assert len(nb) == 2, (case, e, nb)
    tris = []
    seen = set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        prev, cur = None, start
        while True:
            nxt = [n for n in adj[cur] if n != prev]
            # two segments between the same pair of edges cannot occur (a loop has >= 3 edges)
            n = nxt[0] if nxt[0] not in seen or (nxt[0] == start and len(loop) > 2) else nxt[-1]
            if n == start:
                break
            if n in seen:
                n = [x for x in nxt if x not in seen][0]
            loop.append(n)
            seen.add(n)
            prev, cur = cur, n
        assert len(loop) >= 3, (case, loop)
        pts = [edge_mid(e) for e in loop]
        fan = [(0, i, i + 1) for i in range(1, len(loop) - 1)]
        vote = 0.0
        for (a, b, c) in fan:
            nrm = np.cross(pts[b] - pts[a], pts[c] - pts[a])
            cen = (pts[a] + pts[b] + pts[c]) / 3.0
            vote += float(np.dot(nrm, trilinear_grad(case, cen)))
        if vote > 0:            # normal along the gradient = pointing INTO the inside region: flip
            loop = loop[::-1]
            fan = [(0, i, i + 1) for i in range(1, len(loop) - 1)]
        for (a, b, c) in fan:
            tris.append((loop[a], loop[b], loop[c]))
    return tris


def main():
    table = [triangulate(c) for c in range(256)]
    max_t = max(len(t) for t in table)
    flat = np.full((256, max_t * 3), -1, dtype=np.int8)
    for c, t in enumerate(table):
        for i, tri in enumerate(t):
            flat[c, 3 * i:3 * i + 3] = tri
    ntri = [len(t) for t in table]
    hdr = ["// GENERATED by tools/gen_mc_tables.py — marching-cubes case table derived from first principles (see the",
           "// generator for the construction); corner c = cx + 2 cy + 4 cz, edge e = axis * 4 + j.",
           "#pragma once",
           "#define MC_MAX_TRIS %d" % max_t,
           "__constant__ unsigned char c_mc_ntri[256] = {%s};" % ", ".join(str(n) for n in ntri),
           "__constant__ signed char c_mc_tris[256][%d] = {" % (max_t * 3)]
    for c in range(256):
        hdr.append("  {%s}," % ", ".join(str(int(v)) for v in flat[c]))
    hdr.append("};")
    hdr.append("// edge e -> (corner a, corner b)")
    hdr.append("__constant__ unsigned char c_mc_edge_corner[12][2] = {%s};" % ", ".join("{%d, %d}" % e for e in EDGES))
    open(os.path.join(ROOT, "surf_b200", "csrc", "mc_tables.cuh"), "w").write("\n".join(hdr) + "\n")
    py = ['"""GENERATED by tools/gen_mc_tables.py (same table as csrc/mc_tables.cuh), used by the CPU checker in tests."""',
          "MC_MAX_TRIS = %d" % max_t,
          "EDGES = %r" % (EDGES,),
          "TRIS = %r" % ([[tuple(int(x) for x in tri) for tri in t] for t in table],)]
    open(os.path.join(ROOT, "oracle", "mc_tables.py"), "w").write("\n".join(py) + "\n")
    print("max triangles per cube:", max_t, " total table triangles:", sum(ntri))


if __name__ == "__main__":
    main()
