#!/usr/bin/env python
"""Builds oracle/_ref/libref.so from the REFERENCE's own sources -- TEST INFRASTRUCTURE ONLY.

The reference cannot be built with its own build system here (public headers absent, pbvh.c is two
concatenated copies, TBB / OpenSubdiv / bmesh missing).  But the functions on the hot path are plain C: this script
cuts them VERBATIM out of /root/reference at build time (by name, inside a pinned line window, up to the closing brace
in column 0), writes them into one translation unit under oracle/_ref/ (git-ignored: no reference source ever enters
the repository) behind oracle/ref_shim.h (the absent headers' types and macros), and compiles that together with
oracle/ref_api.c (flat-array entry points) into oracle/_ref/libref.so.  tests/test_ref_pin.py then drives the oracle
and libref.so with the same arrays and asserts bit equality: that is what pins the oracle.

    python oracle/ref_extract.py [--reference /root/reference] [--keep-tu]

Without /root/reference (the GPU box) it does nothing and the prebuilt .so is used as it is.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
G = "source/dune/"

# (file, kind, name, first line of the window the definition must start in).  kind: "fn" = a function whose signature
# starts at a line beginning with its return type and ends at the first "}" in column 0; "typedef" = `typedef struct
# NAME {` ... `} NAME;`; "define" = one `#define NAME` line.  pbvh.c holds two concatenated copies; the second, complete one
# starts at line 1913, hence the window.
CHUNKS = [
    (G + "lib/intern/math_base_inline.c", "fn", "min_ff", 380),
    (G + "lib/intern/math_base_inline.c", "fn", "max_ff", 380),
    (G + "lib/intern/math_base_inline.c", "fn", "min_ii", 420),
    (G + "lib/intern/math_base_inline.c", "fn", "max_ii", 420),
    (G + "lib/intern/math_vector_inline.c", "fn", "zero_v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "copy_v3_v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "copy_v4_v4", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "add_v3_v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "sub_v3_v3v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "mul_v3_fl", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "mul_v3_v3fl", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "madd_v3_v3v3fl", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "dot_v3v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "len_squared_v3v3", 1),
    (G + "lib/intern/math_vector_inline.c", "fn", "add_newell_cross_v3_v3v3", 900),
    (G + "lib/intern/math_vector_inline.c", "fn", "normalize_v3_v3_length", 1100),
    (G + "lib/intern/math_vector_inline.c", "fn", "normalize_v3_v3", 1100),
    (G + "lib/intern/math_vector_inline.c", "fn", "normalize_v3", 1100),
    (G + "lib/intern/math_geom.cc", "fn", "normal_tri_v3", 20),
    (G + "lib/intern/math_geom.cc", "fn", "normal_quad_v3", 20),
    (G + "kernel/intern/mesh_evaluate.c", "fn", "mesh_calc_ngon_normal", 30),
    (G + "kernel/intern/mesh_evaluate.c", "fn", "BKE_mesh_calc_poly_normal", 30),
    (G + "kernel/intern/paint.c", "fn", "paint_is_face_hidden", 1200),
    (G + "kernel/intern/paint.c", "fn", "paint_is_grid_face_hidden", 1200),
    (G + "kernel/intern/pbvh.c", "define", "STACK_FIXED_DEPTH", 1913),
    (G + "kernel/intern/pbvh.c", "typedef", "PBVHStack", 1913),
    (G + "kernel/intern/pbvh.c", "typedef", "PBVHIter", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BB_reset", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BB_expand", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BB_expand_with_bb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BB_widest_axis", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BBC_update_centroid", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "update_node_vb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "face_materials_match", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "grid_materials_match", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "partition_indices", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "partition_indices_material", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_grow_nodes", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "map_insert_vert", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "build_mesh_leaf_node", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "update_vb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_count_grid_quads", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "build_grid_leaf_node", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "build_leaf", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "leaf_needs_material_split", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "build_sub", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_build", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_build_mesh", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_build_grids", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_new", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_iter_begin", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_iter_end", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_stack_push", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_iter_next", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_iter_next_occluded", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_search_gather", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "update_search_cb", 1913),
    (G + "kernel/intern/pbvh.c", "typedef", "PBVHUpdateData", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_update_normals_clear_task_cb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_update_normals_accum_task_cb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_update_normals_store_task_cb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_faces_update_normals", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_update_BB_redraw_task_cb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_update_BB_redraw", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_flush_bb", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_update_bounds", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_mark_update", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_mark_rebuild_draw", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_fully_hidden_set", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_vert_mark_update", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_get_verts", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_num_verts", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_get_grids", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_get_BB", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_node_get_original_BB", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "pbvh_vertex_iter_init", 1913),
    (G + "kernel/intern/pbvh.c", "fn", "BKE_pbvh_parallel_range_settings", 1913),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_coord_to_elem", 300),
    (G + "kernel/intern/subdiv_ccg.c", "typedef", "RecalcInnerNormalsTLSData", 600),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_recalc_inner_face_normals", 600),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_average_inner_face_normals", 600),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "average_grid_element_value_v3", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "average_grid_element", 860),
    (G + "kernel/intern/subdiv_ccg.c", "typedef", "GridElementAccumulator", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "element_accumulator_init", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "element_accumulator_add", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "element_accumulator_mul_fl", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "element_accumulator_copy", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_average_inner_face_grids", 860),
    (G + "kernel/intern/subdiv_ccg.c", "typedef", "AverageGridsBoundariesTLSData", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_average_grids_boundary", 860),
    (G + "kernel/intern/subdiv_ccg.c", "fn", "subdiv_ccg_average_grids_corners", 860),
]


def cut(lines, kind, name, first):
    """-> (start, end) 1-based inclusive"""
    if kind == "define":
        for i in range(first - 1, len(lines)):
            if re.match(r"^#define %s\b" % re.escape(name), lines[i]):
                return i + 1, i + 1
        raise SystemExit("#define %s not found" % name)
    if kind == "typedef":
        pat = re.compile(r"^typedef struct %s \{" % re.escape(name))
        for i in range(first - 1, len(lines)):
            if pat.match(lines[i]):
                for j in range(i, len(lines)):
                    if lines[j].startswith("}"):
                        if not re.match(r"^\} %s;" % re.escape(name), lines[j]):
                            raise SystemExit("typedef %s closes with %r" % (name, lines[j]))
                        return i + 1, j + 1
        raise SystemExit("typedef %s not found" % name)
    # a definition: the name followed by "(" on a line that starts in column 0 with a type (not a call, not a prototype)
    pat = re.compile(r"^[A-Za-z_][\w \*]*[ \*]%s\(" % re.escape(name))
    for i in range(first - 1, len(lines)):
        if not pat.match(lines[i]):
            continue
        # prototype? the statement ends with ";" before any "{"
        j = i
        while j < len(lines) and "{" not in lines[j] and ";" not in lines[j]:
            j += 1
        if j < len(lines) and lines[j].rstrip().endswith(";") and "{" not in lines[j]:
            continue
        for k in range(j, len(lines)):
            if lines[k].startswith("}"):
                return i + 1, k + 1
    raise SystemExit("function %s not found" % name)


def main():
    ref = "/root/reference"
    if "--reference" in sys.argv:
        ref = sys.argv[sys.argv.index("--reference") + 1]
    so = os.path.join(OUT, "libref.so")
    if not os.path.isdir(ref):
        print("ref_extract: %s is absent; %s" % (ref, "using the prebuilt " + so if os.path.exists(so) else "no libref.so"))
        return 0
    os.makedirs(OUT, exist_ok=True)
    tu = os.path.join(OUT, "ref_tu.c")
    cache = {}
    parts = ['/* GENERATED by oracle/ref_extract.py from %s -- never committed */\n#include "../ref_shim.h"\n' % ref]
    manifest = []
    for path, kind, name, first in CHUNKS:
        if path not in cache:
            with open(os.path.join(ref, path), errors="replace") as f:
                cache[path] = f.read().split("\n")
        a, b = cut(cache[path], kind, name, first)
        manifest.append("%s:%d-%d %s" % (path, a, b, name))
        parts.append("/* ---- %s:%d-%d ---- */\n#line %d \"%s\"\n%s\n" % (path, a, b, a, path, "\n".join(cache[path][a - 1:b])))
    parts.append('#include "../ref_glue.inc"\n')
    with open(tu, "w") as f:
        f.write("\n".join(parts))
    # the reference's own internal header, next to the unit (git-ignored like it)
    hdr = os.path.join(OUT, "pbvh_intern.h")
    with open(os.path.join(ref, G + "kernel/intern/pbvh_intern.h"), errors="replace") as f:
        open(hdr, "w").write(f.read())
    # what was cut, file:line per function -- names and line ranges only; a copy is committed as oracle/ref_manifest.txt
    for mf in (os.path.join(OUT, "manifest.txt"), os.path.join(HERE, "ref_manifest.txt")):
        with open(mf, "w") as f:
            f.write("\n".join(manifest) + "\n")
    cmd = [GCC, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11", "-w", "-I" + OUT, "-I" + HERE, "-o", so, tu,
           os.path.join(HERE, "ref_api.c"), "-lm"]
    subprocess.run(cmd, check=True)
    if "--keep-tu" not in sys.argv:
        os.remove(tu)  # only the .so travels to the GPU box
        os.remove(hdr)
    print(so)
    return 0


if __name__ == "__main__":
    sys.exit(main())
