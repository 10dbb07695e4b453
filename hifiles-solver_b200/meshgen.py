"""Synthetic structured meshes in the Gambit neutral (.neu) format the reference reads (reference
src/mesh_reader.cpp:105-393), and matching input files.  Used by the tests, smoke() and bench.py: the reference's own
64^3 / 32^3 Taylor-Green meshes are not shipped (testcases/.../.MISSING_LARGE_BLOBS), only the 15^3 one.

Record formats follow the shipped Taylor-Green-Vortex-hex.neu: 6-line header + counts line, `id x y z` node lines,
`id 4 8 n1..n7 \\n n8` brick records (node order -> internal shape slots 0,2,4,6,1,3,5,7, mesh_reader.cpp:240-241),
`id 2 4 n1..n4` quad records (slots 0,1,3,2, :205-206), and boundary sets of `cell type face` triples with Gambit face
ids (hex: 1..6 -> local faces 0,3,5,1,4,2, :336-350; quad: k-1, :333-334)."""
import io
import os

import numpy as np


def _header(name, n_nodes, n_cells, n_bsets, ndim):
    return ("        CONTROL INFO 2.3.16\n** GAMBIT NEUTRAL FILE\n%s\nPROGRAM:                Gambit     VERSION:  2.3.16\n\n"
            "     NUMNP     NELEM     NGRPS    NBSETS     NDFCD     NDFVL\n%10d%10d%10d%10d%10d%10d\nENDOFSECTION\n"
            % (name, n_nodes, n_cells, 1, n_bsets, ndim, ndim))


def _write_rows(f, arr, fmt):
    buf = io.StringIO()
    np.savetxt(buf, arr, fmt=fmt)
    f.write(buf.getvalue())


def _group(f, n_cells):
    f.write("       ELEMENT GROUP 2.3.16\nGROUP:          1 ELEMENTS: %10d MATERIAL:          2 NFLAGS:          1\n"
            "                           fluid\n       0\n" % n_cells)
    ids = np.arange(1, n_cells + 1)
    pad = (-n_cells) % 10
    rows = np.concatenate([ids, np.zeros(pad, dtype=ids.dtype)]).reshape(-1, 10)
    buf = io.StringIO()
    np.savetxt(buf, rows[:-1] if pad else rows, fmt="%8d")
    f.write(buf.getvalue())
    if pad:
        f.write("".join("%8d" % v for v in rows[-1][:10 - pad]) + "\n")
    f.write("ENDOFSECTION\n")


def hex_box(path, n, lengths=(2 * np.pi, 2 * np.pi, 2 * np.pi), bcs=None, origin=(0., 0., 0.), warp=0.0):
    """n = N or (Nx,Ny,Nz) linear bricks on a box.  bcs maps side ('x-','x+','y-','y+','z-','z+') -> boundary-group
    name; default: every side in one group 'Cyclic'.  warp > 0 displaces interior nodes smoothly (non-affine elements;
    boundary nodes stay put so periodic pairing still holds)."""
    if np.isscalar(n):
        n = (n, n, n)
    nx, ny, nz = n
    if bcs is None:
        bcs = {s: "Cyclic" for s in ("x-", "x+", "y-", "y+", "z-", "z+")}
    px, py, pz = nx + 1, ny + 1, nz + 1
    gx, gy, gz = np.meshgrid(np.arange(px), np.arange(py), np.arange(pz), indexing="ij")
    x = origin[0] + lengths[0] * gx / nx
    y = origin[1] + lengths[1] * gy / ny
    z = origin[2] + lengths[2] * gz / nz
    if warp:
        sx, sy, sz = np.sin(np.pi * gx / nx), np.sin(np.pi * gy / ny), np.sin(np.pi * gz / nz)
        bump = warp * sx * sy * sz
        x = x + bump * lengths[0] / nx * np.sin(2 * np.pi * gy / ny + 0.3)
        y = y + bump * lengths[1] / ny * np.sin(2 * np.pi * gz / nz + 0.7)
        z = z + bump * lengths[2] / nz * np.sin(2 * np.pi * gx / nx + 1.1)
    nid = 1 + gx + px * (gy + py * gz)  # node id (1-based), x fastest
    order = np.argsort(nid.ravel())
    nodes = np.column_stack([nid.ravel()[order], x.ravel()[order], y.ravel()[order], z.ravel()[order]])

    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cid = 1 + cx + nx * (cy + ny * cz)  # cell id (1-based), x fastest
    corner = lambda dx, dy, dz: nid[cx + dx, cy + dy, cz + dz]
    # file node m -> internal slot [0,2,4,6,1,3,5,7][m]; slot = ix + 2*iy + 4*iz
    conn = np.stack([cid, np.full_like(cid, 4), np.full_like(cid, 8),
                     corner(0, 0, 0), corner(0, 1, 0), corner(0, 0, 1), corner(0, 1, 1),
                     corner(1, 0, 0), corner(1, 1, 0), corner(1, 0, 1), corner(1, 1, 1)], axis=-1).reshape(-1, 11)
    conn = conn[np.argsort(conn[:, 0])]

    side_faces = {
        "z-": (cid[:, :, 0], 1), "y+": (cid[:, ny - 1, :], 2), "z+": (cid[:, :, nz - 1], 3),
        "y-": (cid[:, 0, :], 4), "x-": (cid[0, :, :], 5), "x+": (cid[nx - 1, :, :], 6)}
    groups = {}
    for side, name in bcs.items():
        cells, k = side_faces[side]
        rows = np.column_stack([np.sort(cells.ravel()), np.full(cells.size, 4), np.full(cells.size, k)])
        groups.setdefault(name, []).append(rows)

    with open(path, "w") as f:
        f.write(_header(os.path.basename(path), nodes.shape[0], conn.shape[0], len(groups), 3))
        f.write("   NODAL COORDINATES 2.3.16\n")
        _write_rows(f, nodes, "%10d %19.16e %19.16e %19.16e")
        f.write("ENDOFSECTION\n      ELEMENTS/CELLS 2.3.16\n")
        buf = io.StringIO()
        np.savetxt(buf, conn, fmt="%8d %2d %2d %8d%8d%8d%8d%8d%8d%8d\n               %8d")
        f.write(buf.getvalue())
        f.write("ENDOFSECTION\n")
        _group(f, conn.shape[0])
        for name, parts in groups.items():
            rows = np.concatenate(parts)
            f.write(" BOUNDARY CONDITIONS 2.3.16\n%32s%8d%8d%8d%8d\n" % (name, 1, rows.shape[0], 0, 6))
            _write_rows(f, rows, "%10d%5d%5d")
            f.write("ENDOFSECTION\n")
    return dict(n_cells=conn.shape[0], n_nodes=nodes.shape[0])


def quad_box(path, n, lengths=(20., 20.), bcs=None, origin=(-10., -10.), warp=0.0):
    """n = N or (Nx,Ny) linear quads.  bcs maps side ('x-','x+','y-','y+') -> boundary-group name (default 'Cyclic')."""
    if np.isscalar(n):
        n = (n, n)
    nx, ny = n
    if bcs is None:
        bcs = {s: "Cyclic" for s in ("x-", "x+", "y-", "y+")}
    px, py = nx + 1, ny + 1
    gx, gy = np.meshgrid(np.arange(px), np.arange(py), indexing="ij")
    x = origin[0] + lengths[0] * gx / nx
    y = origin[1] + lengths[1] * gy / ny
    if warp:
        bump = warp * np.sin(np.pi * gx / nx) * np.sin(np.pi * gy / ny)
        x = x + bump * lengths[0] / nx * np.sin(2 * np.pi * gy / ny + 0.3)
        y = y + bump * lengths[1] / ny * np.sin(2 * np.pi * gx / nx + 0.7)
    nid = 1 + gx + px * gy
    order = np.argsort(nid.ravel())
    nodes = np.column_stack([nid.ravel()[order], x.ravel()[order], y.ravel()[order]])
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    cid = 1 + cx + nx * cy
    corner = lambda dx, dy: nid[cx + dx, cy + dy]
    conn = np.stack([cid, np.full_like(cid, 2), np.full_like(cid, 4), corner(0, 0), corner(1, 0), corner(1, 1), corner(0, 1)], axis=-1).reshape(-1, 7)
    conn = conn[np.argsort(conn[:, 0])]
    side_faces = {"y-": (cid[:, 0], 1), "x+": (cid[nx - 1, :], 2), "y+": (cid[:, ny - 1], 3), "x-": (cid[0, :], 4)}
    groups = {}
    for side, name in bcs.items():
        cells, k = side_faces[side]
        rows = np.column_stack([np.sort(cells.ravel()), np.full(cells.size, 2), np.full(cells.size, k)])
        groups.setdefault(name, []).append(rows)
    with open(path, "w") as f:
        f.write(_header(os.path.basename(path), nodes.shape[0], conn.shape[0], len(groups), 2))
        f.write("   NODAL COORDINATES 2.3.16\n")
        _write_rows(f, nodes, "%10d %19.16e %19.16e")
        f.write("ENDOFSECTION\n      ELEMENTS/CELLS 2.3.16\n")
        _write_rows(f, conn, "%8d %2d %2d %8d%8d%8d%8d")
        f.write("ENDOFSECTION\n")
        _group(f, conn.shape[0])
        for name, parts in groups.items():
            rows = np.concatenate(parts)
            f.write(" BOUNDARY CONDITIONS 2.3.16\n%32s%8d%8d%8d%8d\n" % (name, 1, rows.shape[0], 0, 6))
            _write_rows(f, rows, "%10d%5d%5d")
            f.write("ENDOFSECTION\n")
    return dict(n_cells=conn.shape[0], n_nodes=nodes.shape[0])


def _write_neu(path, ndim, nodes, cells, groups):
    """nodes: (id, x, y[, z]) rows; cells: list of (id, ntype, [node ids]) ; groups: {name: rows (cell, ntype, face)}"""
    with open(path, "w") as f:
        f.write(_header(os.path.basename(path), nodes.shape[0], len(cells), len(groups), ndim))
        f.write("   NODAL COORDINATES 2.3.16\n")
        _write_rows(f, nodes, "%10d " + " ".join(["%19.16e"] * ndim))
        f.write("ENDOFSECTION\n      ELEMENTS/CELLS 2.3.16\n")
        out = []
        for cid, ntype, vs in cells:
            head = "%8d %2d %2d " % (cid, ntype, len(vs))
            first, rest = vs[:7], vs[7:]
            out.append(head + "".join("%8d" % v for v in first) + "\n")
            while rest:
                out.append("               " + "".join("%8d" % v for v in rest[:7]) + "\n")
                rest = rest[7:]
        f.write("".join(out))
        f.write("ENDOFSECTION\n")
        _group(f, len(cells))
        for name, rows in groups.items():
            rows = np.asarray(rows)
            f.write(" BOUNDARY CONDITIONS 2.3.16\n%32s%8d%8d%8d%8d\n" % (name, 1, rows.shape[0], 0, 6))
            _write_rows(f, rows, "%10d%5d%5d")
            f.write("ENDOFSECTION\n")
    xyz = {int(r[0]): r[1:] for r in nodes}
    cent = np.array([np.mean([xyz[v] for v in vs], axis=0) for _, _, vs in cells])
    return dict(n_cells=len(cells), n_nodes=nodes.shape[0], centroids=cent)


def slab_partition(centroids, nproc, axis=0):
    """part[global cell] for any mesh: nproc slabs of (nearly) equal cell count along one axis, from the cell centroids a
    generator returned.  A stand-in for the reference's ParMETIS call on the generated test meshes."""
    order = np.argsort(centroids[:, axis], kind="stable")
    part = np.zeros(len(order), dtype=np.int32)
    part[order] = (np.arange(len(order)) * nproc) // len(order)
    return part


def quad8_box(path, n, lengths=(20., 20.), bcs=None, origin=(-10., -10.), curve=0.0):
    """n = N or (Nx,Ny) eight-node (serendipity) quadrilaterals: `id 2 8 c0 m01 c1 m12 c2 m23 c3 m30`, corners and mid-edge
    nodes alternating counter-clockwise as Gambit writes them (mesh_reader.cpp:199-206 maps them to corners 0-3, mid-edge
    nodes 4-7).  curve > 0 bows the interior edges: mid-edge nodes displaced normal to their edge."""
    if np.isscalar(n):
        n = (n, n)
    nx, ny = n
    if bcs is None:
        bcs = {s: "Cyclic" for s in ("x-", "x+", "y-", "y+")}
    px, py = 2 * nx + 1, 2 * ny + 1
    gi, gj = np.meshgrid(np.arange(px), np.arange(py), indexing="ij")
    used = ~((gi % 2 == 1) & (gj % 2 == 1))
    x = origin[0] + lengths[0] * gi / (2. * nx)
    y = origin[1] + lengths[1] * gj / (2. * ny)
    if curve:
        interior = (gi > 0) & (gi < px - 1) & (gj > 0) & (gj < py - 1)
        on_x_edge = (gi % 2 == 1) & (gj % 2 == 0) & interior   # mid node of an edge along x: move in y
        on_y_edge = (gi % 2 == 0) & (gj % 2 == 1) & interior
        y = y + np.where(on_x_edge, curve * lengths[1] / ny * np.sin(1.3 * gi + 0.4 * gj), 0.)
        x = x + np.where(on_y_edge, curve * lengths[0] / nx * np.cos(0.7 * gi + 1.1 * gj), 0.)
    nid = np.zeros((px, py), dtype=np.int64)
    order = np.argsort((gi + px * gj)[used])
    ids = np.empty(order.size, dtype=np.int64)
    ids[order] = 1 + np.arange(order.size)
    nid[used] = ids
    sel = np.argsort(nid[used])
    nodes = np.column_stack([nid[used][sel], x[used][sel], y[used][sel]])
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    cid = 1 + cx + nx * cy
    at = lambda di, dj: nid[2 * cx + di, 2 * cy + dj]
    conn = np.stack([cid, np.full_like(cid, 2), np.full_like(cid, 8), at(0, 0), at(1, 0), at(2, 0), at(2, 1), at(2, 2), at(1, 2), at(0, 2), at(0, 1)],
                    axis=-1).reshape(-1, 11)
    conn = conn[np.argsort(conn[:, 0])]
    side_faces = {"y-": (cid[:, 0], 1), "x+": (cid[nx - 1, :], 2), "y+": (cid[:, ny - 1], 3), "x-": (cid[0, :], 4)}
    groups = {}
    for side, name in bcs.items():
        cells, k = side_faces[side]
        rows = np.column_stack([np.sort(cells.ravel()), np.full(cells.size, 2), np.full(cells.size, k)])
        groups.setdefault(name, []).append(rows)
    with open(path, "w") as f:
        f.write(_header(os.path.basename(path), nodes.shape[0], conn.shape[0], len(groups), 2))
        f.write("   NODAL COORDINATES 2.3.16\n")
        _write_rows(f, nodes, "%10d %19.16e %19.16e")
        f.write("ENDOFSECTION\n      ELEMENTS/CELLS 2.3.16\n")
        buf = io.StringIO()
        np.savetxt(buf, conn, fmt="%8d %2d %2d %8d%8d%8d%8d%8d%8d%8d\n               %8d")
        f.write(buf.getvalue())
        f.write("ENDOFSECTION\n")
        _group(f, conn.shape[0])
        for name, parts in groups.items():
            rows = np.concatenate(parts)
            f.write(" BOUNDARY CONDITIONS 2.3.16\n%32s%8d%8d%8d%8d\n" % (name, 1, rows.shape[0], 0, 6))
            _write_rows(f, rows, "%10d%5d%5d")
            f.write("ENDOFSECTION\n")
    return dict(n_cells=conn.shape[0], n_nodes=nodes.shape[0])


def hex20_box(path, n, lengths=(2 * np.pi,) * 3, bcs=None, origin=(0., 0., 0.), warp=0.0):
    """n = N or (Nx,Ny,Nz) twenty-node (serendipity) hexahedra: `id 4 20 ...` with the nodes in Gambit's order, the 3x3x3
    lattice of the cell without face and body centres, x slowest, then z, y fastest (mesh_reader.cpp:242-243 sends them to
    corners 0-7 and mid-edge nodes 8-19).  warp > 0 displaces all interior nodes smoothly, so edges become parabolas."""
    if np.isscalar(n):
        n = (n, n, n)
    nx, ny, nz = n
    if bcs is None:
        bcs = {s: "Cyclic" for s in ("x-", "x+", "y-", "y+", "z-", "z+")}
    px, py, pz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    gx, gy, gz = np.meshgrid(np.arange(px), np.arange(py), np.arange(pz), indexing="ij")
    used = ((gx % 2) + (gy % 2) + (gz % 2)) <= 1
    x = origin[0] + lengths[0] * gx / (2. * nx)
    y = origin[1] + lengths[1] * gy / (2. * ny)
    z = origin[2] + lengths[2] * gz / (2. * nz)
    if warp:
        bump = warp * np.sin(np.pi * gx / (2. * nx)) * np.sin(np.pi * gy / (2. * ny)) * np.sin(np.pi * gz / (2. * nz))
        x = x + bump * lengths[0] / nx * np.sin(np.pi * gy / ny + 0.3)
        y = y + bump * lengths[1] / ny * np.sin(np.pi * gz / nz + 0.7)
        z = z + bump * lengths[2] / nz * np.sin(np.pi * gx / nx + 1.1)
    nid = np.zeros((px, py, pz), dtype=np.int64)
    key = (gx + px * (gy + py * gz))[used]
    ids = np.empty(key.size, dtype=np.int64)
    ids[np.argsort(key)] = 1 + np.arange(key.size)
    nid[used] = ids
    sel = np.argsort(ids)
    nodes = np.column_stack([ids[sel], x[used][sel], y[used][sel], z[used][sel]])
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cid = 1 + cx + nx * (cy + ny * cz)
    cols = [cid, np.full_like(cid, 4), np.full_like(cid, 20)]
    for di in range(3):
        for dk in range(3):
            for dj in range(3):
                if (di % 2) + (dj % 2) + (dk % 2) <= 1:
                    cols.append(nid[2 * cx + di, 2 * cy + dj, 2 * cz + dk])
    conn = np.stack(cols, axis=-1).reshape(-1, 23)
    conn = conn[np.argsort(conn[:, 0])]
    side_faces = {
        "z-": (cid[:, :, 0], 1), "y+": (cid[:, ny - 1, :], 2), "z+": (cid[:, :, nz - 1], 3),
        "y-": (cid[:, 0, :], 4), "x-": (cid[0, :, :], 5), "x+": (cid[nx - 1, :, :], 6)}
    groups = {}
    for side, name in bcs.items():
        cells, k = side_faces[side]
        rows = np.column_stack([np.sort(cells.ravel()), np.full(cells.size, 4), np.full(cells.size, k)])
        groups.setdefault(name, []).append(rows)
    with open(path, "w") as f:
        f.write(_header(os.path.basename(path), nodes.shape[0], conn.shape[0], len(groups), 3))
        f.write("   NODAL COORDINATES 2.3.16\n")
        _write_rows(f, nodes, "%10d %19.16e %19.16e %19.16e")
        f.write("ENDOFSECTION\n      ELEMENTS/CELLS 2.3.16\n")
        buf = io.StringIO()
        pad = "\n               "
        np.savetxt(buf, conn, fmt="%8d %2d %2d " + "%8d" * 7 + pad + "%8d" * 7 + pad + "%8d" * 6)
        f.write(buf.getvalue())
        f.write("ENDOFSECTION\n")
        _group(f, conn.shape[0])
        for name, parts in groups.items():
            rows = np.concatenate(parts)
            f.write(" BOUNDARY CONDITIONS 2.3.16\n%32s%8d%8d%8d%8d\n" % (name, 1, rows.shape[0], 0, 6))
            _write_rows(f, rows, "%10d%5d%5d")
            f.write("ENDOFSECTION\n")
    return dict(n_cells=conn.shape[0], n_nodes=nodes.shape[0])


def mixed_box_2d(path, n, lengths=(20., 20.), bcs=None, origin=(-10., -10.), kind="tri", warp=0.0, curve=0.0):
    """Nx x Ny cells on a rectangle.  kind: 'tri' (every cell split into two triangles along its (0,0)-(1,1) diagonal),
    'quad', or 'mixed' (quads in the left half, triangles in the right half: BASELINE config 2's element mix).
    'tri6' = quadratic six-node triangles (`id 3 6 v0 m01 v1 m12 v2 m20`, mesh_reader.cpp:192-197); curve > 0 bows the
    interior edges (mid-edge nodes displaced normal to their edge), like the shipped cylinder mesh's curved elements.
    Gambit records: triangles `id 3 3 n1 n2 n3` (counter-clockwise; edge k joins nodes k, k+1: mesh_reader.cpp:192-197,
    :333-334), quads as in quad_box."""
    if np.isscalar(n):
        n = (n, n)
    nx, ny = n
    if bcs is None:
        bcs = {s: "Cyclic" for s in ("x-", "x+", "y-", "y+")}
    px, py = nx + 1, ny + 1
    gx, gy = np.meshgrid(np.arange(px), np.arange(py), indexing="ij")
    x = origin[0] + lengths[0] * gx / nx
    y = origin[1] + lengths[1] * gy / ny
    if warp:
        bump = warp * np.sin(np.pi * gx / nx) * np.sin(np.pi * gy / ny)
        x = x + bump * lengths[0] / nx * np.sin(2 * np.pi * gy / ny + 0.3)
        y = y + bump * lengths[1] / ny * np.sin(2 * np.pi * gx / nx + 0.7)
    nid = 1 + gx + px * gy
    order = np.argsort(nid.ravel())
    nodes = np.column_stack([nid.ravel()[order], x.ravel()[order], y.ravel()[order]])
    cells, groups = [], {}
    xy = {int(nid[a, b]): (x[a, b], y[a, b]) for a in range(px) for b in range(py)}
    mids, extra = {}, []

    def mid(a, b, interior):
        key = (min(a, b), max(a, b))
        if key not in mids:
            (xa, ya), (xb, yb) = xy[key[0]], xy[key[1]]
            xm, ym = 0.5 * (xa + xb), 0.5 * (ya + yb)
            if curve and interior:
                xm, ym = xm - curve * (yb - ya), ym + curve * (xb - xa)
            mids[key] = px * py + 1 + len(extra)
            extra.append((mids[key], xm, ym))
        return mids[key]

    def bface(side, cid, ntype, k):
        if side in bcs:
            groups.setdefault(bcs[side], []).append((cid, ntype, k))

    cid = 0
    for j in range(ny):
        for i in range(nx):
            v00, v10, v11, v01 = nid[i, j], nid[i + 1, j], nid[i + 1, j + 1], nid[i, j + 1]
            if kind == "tri6":
                for tri, sides in (((v00, v10, v11), (j == 0 and "y-", i == nx - 1 and "x+", False)),
                                   ((v00, v11, v01), (False, j == ny - 1 and "y+", i == 0 and "x-"))):
                    cid += 1
                    on_bdy = [(j == 0, i == nx - 1, False), (False, j == ny - 1, i == 0)][0 if tri[1] == v10 else 1]
                    ms = [mid(tri[q], tri[(q + 1) % 3], not on_bdy[q]) for q in range(3)]
                    cells.append((cid, 3, [tri[0], ms[0], tri[1], ms[1], tri[2], ms[2]]))
                    for q, side in enumerate(sides):
                        if side: bface(side, cid, 3, q + 1)
                continue
            as_quad = kind == "quad" or (kind == "mixed" and i < nx // 2)
            if as_quad:
                cid += 1
                cells.append((cid, 2, [v00, v10, v11, v01]))
                if j == 0: bface("y-", cid, 2, 1)
                if i == nx - 1: bface("x+", cid, 2, 2)
                if j == ny - 1: bface("y+", cid, 2, 3)
                if i == 0: bface("x-", cid, 2, 4)
            else:
                cid += 1
                cells.append((cid, 3, [v00, v10, v11]))
                if j == 0: bface("y-", cid, 3, 1)
                if i == nx - 1: bface("x+", cid, 3, 2)
                cid += 1
                cells.append((cid, 3, [v00, v11, v01]))
                if j == ny - 1: bface("y+", cid, 3, 2)
                if i == 0: bface("x-", cid, 3, 3)
    if extra:
        nodes = np.vstack([nodes, np.array(extra)])
    return _write_neu(path, 2, nodes, cells, groups)


def tri_box(path, n, **kw):
    return mixed_box_2d(path, n, kind="tri", **kw)


_KUHN = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
# Gambit face number of the local faces (reference src/mesh_reader.cpp:336-350, inverted)
_TET_GAMBIT_FACE = {3: 1, 2: 2, 0: 3, 1: 4}          # local face f is opposite local vertex f
_PRI_GAMBIT_FACE = {2: 1, 3: 2, 4: 3, 0: 4, 1: 5}    # local faces: 0 bottom, 1 top triangle, 2..4 the sides over edges 01, 12, 20
_HEX_GAMBIT_FACE = {"z-": 1, "y+": 2, "z+": 3, "y-": 4, "x-": 5, "x+": 6}


def mixed_box_3d(path, n, lengths=(2 * np.pi,) * 3, bcs=None, origin=(0., 0., 0.), kind="tet", warp=0.0):
    """Nx x Ny x Nz cubes on a box.  kind:
      'tet'      every cube split into six tetrahedra around its (0,0,0)-(1,1,1) diagonal (Kuhn), which cuts every cube
                 face along the diagonal from its lowest to its highest corner, so the split is conforming and periodic;
      'pri'      every cube split into two prisms extruded along y, triangles in the x-z plane cut along the same diagonal;
      'hex'      bricks;
      'hexpri'   bricks for z-index < Nz/2, prisms above (they meet through the prisms' quadrilateral z faces);
      'pritet'   prisms for y-index < Ny/2, tetrahedra above (they meet through triangles in x-z planes).
    Records: tets `id 6 4 n1..n4` (slots 0..3), prisms `id 5 6 n1..n6` (bottom triangle then top triangle), bricks as in
    hex_box (mesh_reader.cpp:207-246)."""
    if np.isscalar(n):
        n = (n, n, n)
    nx, ny, nz = n
    sides = ("x-", "x+", "y-", "y+", "z-", "z+")
    if bcs is None:
        bcs = {s: "Cyclic" for s in sides}
    px, py, pz = nx + 1, ny + 1, nz + 1
    gx, gy, gz = np.meshgrid(np.arange(px), np.arange(py), np.arange(pz), indexing="ij")
    x = origin[0] + lengths[0] * gx / nx
    y = origin[1] + lengths[1] * gy / ny
    z = origin[2] + lengths[2] * gz / nz
    if warp:
        bump = warp * np.sin(np.pi * gx / nx) * np.sin(np.pi * gy / ny) * np.sin(np.pi * gz / nz)
        x = x + bump * lengths[0] / nx * np.sin(2 * np.pi * gy / ny + 0.3)
        y = y + bump * lengths[1] / ny * np.sin(2 * np.pi * gz / nz + 0.7)
        z = z + bump * lengths[2] / nz * np.sin(2 * np.pi * gx / nx + 1.1)
    nid = 1 + gx + px * (gy + py * gz)
    order = np.argsort(nid.ravel())
    nodes = np.column_stack([nid.ravel()[order], x.ravel()[order], y.ravel()[order], z.ravel()[order]])
    cells, groups = [], {}

    def on_side(ijk_list):
        """box side all the given lattice points lie on, or None"""
        a = np.array(ijk_list)
        for d, (lo, hi) in enumerate((("x-", "x+"), ("y-", "y+"), ("z-", "z+"))):
            if np.all(a[:, d] == 0): return lo
            if np.all(a[:, d] == n[d]): return hi
        return None

    def bface(side, cid, ntype, k):
        if side is not None and side in bcs:
            groups.setdefault(bcs[side], []).append((cid, ntype, k))

    tet_faces = [(1, 2, 3), (0, 3, 2), (0, 1, 3), (0, 2, 1)]
    pri_faces = [(0, 2, 1), (3, 4, 5), (0, 1, 4, 3), (1, 2, 5, 4), (2, 0, 3, 5)]
    cid = 0
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                if kind == "hexpri": sub = "hex" if k < nz // 2 else "pri"
                elif kind == "pritet": sub = "pri" if j < ny // 2 else "tet"
                else: sub = kind
                P = lambda a, b, c: (i + a, j + b, k + c)
                if sub == "hex":
                    cid += 1
                    pts = [P(0, 0, 0), P(0, 1, 0), P(0, 0, 1), P(0, 1, 1), P(1, 0, 0), P(1, 1, 0), P(1, 0, 1), P(1, 1, 1)]
                    cells.append((cid, 4, [nid[q] for q in pts]))
                    for side, cond in (("z-", k == 0), ("y+", j == ny - 1), ("z+", k == nz - 1), ("y-", j == 0), ("x-", i == 0), ("x+", i == nx - 1)):
                        if cond: bface(side, cid, 4, _HEX_GAMBIT_FACE[side])
                elif sub == "pri":
                    # triangles in the x-z plane (counter-clockwise seen from -y, so that bottom -> top points along +y)
                    for tri in ([P(0, 0, 0), P(1, 0, 1), P(1, 0, 0)], [P(0, 0, 0), P(0, 0, 1), P(1, 0, 1)]):
                        cid += 1
                        pts = tri + [(a, b + 1, c) for (a, b, c) in tri]
                        cells.append((cid, 5, [nid[q] for q in pts]))
                        for f, vs in enumerate(pri_faces):
                            bface(on_side([pts[v] for v in vs]), cid, 5, _PRI_GAMBIT_FACE[f])
                else:
                    for perm in _KUHN:
                        cur = [0, 0, 0]
                        pts = [P(*cur)]
                        for ax in perm:
                            cur[ax] = 1
                            pts.append(P(*cur))
                        # positive orientation: even permutations as they are, odd ones with two vertices swapped
                        sign = np.linalg.det(np.array([np.subtract(pts[m], pts[0]) for m in (1, 2, 3)], dtype=float))
                        if sign < 0:
                            pts[2], pts[3] = pts[3], pts[2]
                        cid += 1
                        cells.append((cid, 6, [nid[q] for q in pts]))
                        for f, vs in enumerate(tet_faces):
                            bface(on_side([pts[v] for v in vs]), cid, 6, _TET_GAMBIT_FACE[f])
    return _write_neu(path, 3, nodes, cells, groups)


# ---- input files -------------------------------------------------------------------------------------------------------
_TGV_DEFAULTS = dict(
    equation=0, viscous=1, riemann_solve_type=3, vis_riemann_solve_type=0, ic_form=7, test_case=0, order=4, dt_type=0,
    dt=1.0e-5, n_steps=1, adv_type=2, LES=0, over_int=0, restart_flag=0, dx_cyclic=6.2831853071795862,
    dy_cyclic=6.2831853071795862, dz_cyclic=6.2831853071795862, p_res=2, write_type=0, monitor_res_freq=1000000,
    plot_freq=1000000, restart_dump_freq=1000000, data_file_name="run", res_norm_type=1, error_norm_type=1, res_norm_field=0,
    upts_type_hexa=0, vcjh_scheme_hexa=1, eta_hexa=0., sparse_hexa=0, upts_type_quad=0, vcjh_scheme_quad=1, eta_quad=0.,
    sparse_quad=0, upts_type_tri=0, fpts_type_tri=0, vcjh_scheme_tri=1, c_tri=0.0, sparse_tri=0, upts_type_tet=0,
    fpts_type_tet=0, vcjh_scheme_tet=1, eta_tet=0.0, sparse_tet=0, upts_type_pri_tri=0, upts_type_pri_1d=0,
    vcjh_scheme_pri_1d=1, eta_pri=0.0, sparse_pri=0, bc_Cyclic_type="cyclic", gamma=1.4, prandtl=0.72, S_gas=120., T_gas=291.15, R_gas=286.9, mu_gas=1.827e-5, fix_vis=1,
    Mach_free_stream=0.1, L_free_stream=1.0, T_free_stream=300., rho_free_stream=0.0008421095852102401,
    Mach_c_ic=0.1, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=0.0008421095852102401, ldg_beta=0.5, ldg_tau=0.)


def write_input(path, mesh_file, **overrides):
    """Input file in the reference's `key value` format (reference src/input.cpp:62-327).  Defaults are the shipped
    Taylor-Green input (testcases/navier-stokes/Taylor_Green_vortex/input_TGV_SD_hex) with BASELINE config 3's changes
    (order 4, adv_type 2); overrides replace or add keys.  A value of None removes a key."""
    opts = dict(_TGV_DEFAULTS)
    opts["mesh_file"] = mesh_file
    opts.update(overrides)
    with open(path, "w") as f:
        for k, v in opts.items():
            if v is None:
                continue
            if isinstance(v, float):
                f.write("%s %.17g\n" % (k, v))
            else:
                f.write("%s %s\n" % (k, v))
    return path


def block_partition(n, blocks):
    """part[global cell] for an n^3 (or (nx,ny,nz)) hex_box split into blocks=(bx,by,bz) equal bricks: the partition a
    k-way graph partitioner converges to on a uniform periodic cube."""
    if np.isscalar(n):
        n = (n, n, n)
    cx, cy, cz = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    cid = cx + n[0] * (cy + n[1] * cz)
    r = (cx * blocks[0] // n[0]) + blocks[0] * ((cy * blocks[1] // n[1]) + blocks[1] * (cz * blocks[2] // n[2]))
    part = np.zeros(cid.size, dtype=np.int32)
    part[cid.ravel()] = r.ravel()
    return part


def blocks_for(nproc):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nproc]
