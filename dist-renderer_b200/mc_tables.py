"""Case tables of the marching-cubes kernel (`csrc/mesh.cu`), derived rather than typed in.

The reference meshes its SDF grid with `skimage.measure.marching_cubes_lewiner` (`core/evaluation/create_mesh.py:164`),
a third-party routine that is not vendored there and not installed here.  What this module fixes is the published
algorithm -- one vertex per sign-changing cube edge, triangles from a 256-row case table -- with the table built by
tracing the iso-contour over the cube's faces, so that every choice is stated once and can be checked case by case
(`tests/test_mesh_cpu.py`) instead of trusting 4 096 hand-copied numbers:

  corners   c = 4*d0 + 2*d1 + d2 with d_k the offset along array axis k (axis 2 fastest, like the volume);
            a corner is *inside* when value < level; case index = sum of 2^c over inside corners.
  edges     e = 4*a + 2*u + v: the edge along axis a whose offsets along the two other axes (in increasing axis order)
            are u and v; it runs from the corner with d_a = 0 (its *owner*) to the one with d_a = 1.
  faces     a face holding two crossing edges joins them; a face holding four (inside corners on one diagonal) joins
            the two edges that meet at each inside corner, i.e. inside corners are cut off one by one.  The rule reads
            only the face's own four corner signs, so the two cubes sharing a face always agree: the mesh has no cracks.
  loops     the face segments form closed loops v_0 .. v_{n-1}; each starts at its lowest edge and runs so that triangle
            normals (right-hand rule) point from inside to outside; loops are emitted in order of their first edge.
  triangles a loop is cut by `split(0, n-1)`, where split(i, j) takes the smallest apex k in (i, j) for which the chords
            (v_i, v_k) and (v_k, v_j) do not lie in a cube face and both sides can be split in turn, and emits
            split(i, k), (v_i, v_k, v_j), split(k, j).  A chord in a face plane (two vertices of one ambiguous face that
            its segments do not join) would be produced by the neighbouring cube as well and leave an edge with four
            triangles; all 256 cases admit a chord-free cut.

Differences from the Lewiner variant are confined to cubes with an ambiguous face or interior (rare on a smooth SDF):
there Lewiner's extra tests may pick the other diagonal or add a centre vertex.  Vertices elsewhere are identical
(linear interpolation on the same edges).
"""
import os

MAX_TRIS = None          # set below (longest row of the table)


def corner_offset(c):
    return ((c >> 2) & 1, (c >> 1) & 1, c & 1)


def corner_index(d):
    return (d[0] << 2) | (d[1] << 1) | d[2]


def edge_corners(e):
    """(owner corner, far corner) of edge e."""
    a, u, v = e >> 2, (e >> 1) & 1, e & 1
    others = [k for k in range(3) if k != a]
    d = [0, 0, 0]
    d[others[0]], d[others[1]] = u, v
    c0 = corner_index(d)
    d[a] = 1
    return c0, corner_index(d)


def edge_mid(e):
    c0, c1 = edge_corners(e)
    p0, p1 = corner_offset(c0), corner_offset(c1)
    return tuple((p0[k] + p1[k]) * 0.5 for k in range(3))


def face_edges(f, s):
    """The four edges lying in the face `axis f, side s`."""
    out = []
    for e in range(12):
        if (e >> 2) == f:
            continue
        c0, _ = edge_corners(e)
        if corner_offset(c0)[f] == s:
            out.append(e)
    return out


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _inside_end(e, case):
    c0, c1 = edge_corners(e)
    return c0 if (case >> c0) & 1 else c1


def case_segments(case):
    """{edge: [(neighbour edge, face axis, face side), (…)]} for the crossing edges of `case`."""
    crossing = [e for e in range(12) if ((case >> edge_corners(e)[0]) ^ (case >> edge_corners(e)[1])) & 1]
    adj = {e: [] for e in crossing}
    for f in range(3):
        for s in range(2):
            on = [e for e in face_edges(f, s) if e in adj]
            if len(on) == 2:
                pairs = [(on[0], on[1])]
            elif len(on) == 4:
                pairs = []
                for e in on:
                    mate = [x for x in on if x != e and _inside_end(x, case) == _inside_end(e, case)]
                    assert len(mate) == 1
                    if e < mate[0]:
                        pairs.append((e, mate[0]))
            else:
                assert len(on) == 0
                pairs = []
            for a, b in pairs:
                adj[a].append((b, f, s))
                adj[b].append((a, f, s))
    assert all(len(v) == 2 for v in adj.values())
    return adj


def case_triangles(case):
    """Triangles (triples of edge ids) of one cube case, in emission order."""
    adj = case_segments(case)
    todo = set(adj)
    tris = []
    while todo:
        start = min(todo)
        ci = corner_offset(_inside_end(start, case))
        mid = edge_mid(start)
        m = tuple(ci[k] - mid[k] for k in range(3))
        first = None
        for b, f, s in adj[start]:
            n = [0, 0, 0]
            n[f] = 1 if s else -1
            mb = edge_mid(b)
            sdir = tuple(mb[k] - mid[k] for k in range(3))
            x = _cross(m, sdir)
            if x[0] * n[0] + x[1] * n[1] + x[2] * n[2] > 0:
                assert first is None
                first = b
        assert first is not None
        loop, prev, cur = [start], start, first
        while cur != start:
            loop.append(cur)
            a, b = adj[cur][0][0], adj[cur][1][0]
            assert a != b and prev in (a, b)
            prev, cur = cur, (b if a == prev else a)
        todo -= set(loop)
        cut = _split(loop, 0, len(loop) - 1)
        assert cut is not None
        tris.extend(cut)
    return tris


def _in_face(a, b):
    return any(a in face_edges(f, s) and b in face_edges(f, s) for f in range(3) for s in range(2))


def _split(loop, i, j):
    """Triangles of the sub-polygon v_i .. v_j (closed by the edge or accepted chord v_i v_j); None if impossible."""
    if j - i < 2:
        return []
    n = len(loop)

    def ok(a, b):
        return b - a == 1 or (a == 0 and b == n - 1) or not _in_face(loop[a], loop[b])
    for k in range(i + 1, j):
        if ok(i, k) and ok(k, j):
            left, right = _split(loop, i, k), _split(loop, k, j)
            if left is not None and right is not None:
                return left + [(loop[i], loop[k], loop[j])] + right
    return None


def tables():
    """(n_tris[256], tri_edges[256][3*MAX_TRIS] padded with -1)."""
    rows = [case_triangles(c) for c in range(256)]
    width = max(len(r) for r in rows)
    n = [len(r) for r in rows]
    flat = []
    for r in rows:
        row = [e for t in r for e in t]
        flat.append(row + [-1] * (3 * width - len(row)))
    return n, flat, width


_N, _T, MAX_TRIS = tables()


def as_header():
    """Text of csrc/mc_tables.inc."""
    out = ["// generated by dist-renderer_b200/mc_tables.py (written by build.py before nvcc runs)",
           "#define DIST_MC_MAX_TRIS %d" % MAX_TRIS,
           "__constant__ signed char c_mc_ntri[256] = {" + ",".join(str(x) for x in _N) + "};",
           "__constant__ signed char c_mc_tri[256][%d] = {" % (3 * MAX_TRIS)]
    for row in _T:
        out.append("  {" + ",".join(str(x) for x in row) + "},")
    out.append("};")
    return "\n".join(out) + "\n"


def write_header(path):
    text = as_header()
    if not os.path.isfile(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


def n_tris():
    return list(_N)


def tri_edges():
    return [list(r) for r in _T]
