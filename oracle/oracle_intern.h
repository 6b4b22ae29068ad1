/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Internal structs. */
#ifndef DUNE_ORACLE_INTERN_H
#define DUNE_ORACLE_INTERN_H

#include "oracle.h"

/* kernel/intern/pbvh_intern.h:4-11 */
typedef struct OrBBC {
  float bmin[3], bmax[3], bcentroid[3];
} OrBBC;

/* kernel/intern/pbvh_intern.h:15-90 (fields on the path only) */
typedef struct OrNode {
  OrBB vb, orig_vb;
  int children_offset;
  int prim_offset; /* prim_indices = pbvh->prim_indices + prim_offset */
  int totprim;
  int *vert_indices;
  int uniq_verts, face_verts;
  int (*face_vert_indices)[3];
  unsigned flag;
  float tmin; /* ray-cast: entry distance of the ray into vb (pbvh_intern.h:81) */
} OrNode;

/* kernel/intern/pbvh_intern.h:98-162 plus the sculpt-session state the brushes need */
struct OrPbvh {
  OrNode *nodes;
  int node_mem_count, totnode;
  int *prim_indices;
  int totprim, totvert, leaf_limit;

  float (*co)[3];  /* MVert.co */
  float (*no)[3];  /* vert_normals */
  float *mask;     /* CD_PAINT_MASK or NULL */
  int totpoly, totloop;
  int *poly_start, *poly_len, *loop_v;
  int (*tri_loop)[3]; /* MLoopTri.tri */
  int (*tri_v)[3];    /* mloop[tri].v resolved */
  int *tri_poly;
  unsigned char *vert_bitmap; /* one byte per vertex */
  /* ordered threaded normals (or_set_ordered_normals): vertex -> incident looptri positions, ascending; position -> leaf */
  int64_t *vt_off;
  int *vt_pos, *pos_node;
  /* material / visibility inputs of the build and the vertex iterator (NULL: one material, nothing hidden) */
  short *poly_mat, *grid_mat;           /* MPoly.mat_nr, DMFlagMat.mat_nr */
  unsigned char *poly_flag, *grid_flag; /* MPoly.flag, DMFlagMat.flag (ME_SMOOTH = 1) */
  unsigned char *vert_flag;             /* MVert.flag (ME_HIDE = 16) */
  unsigned char *grid_hidden;           /* [totgrid * grid_size^2] grid_hidden bit of the element */

  /* sculpt session DAGGER */
  int *nb_off, *nb_idx; /* vertex neighbours, reference order */
  unsigned char *boundary;
  float *automask;
  float (*orig_co)[3], (*orig_no)[3];
  unsigned char *touched; /* per node: undo node pushed this stroke */
  float curve_table[257];
  int has_curve_table;

  int *last_hits, last_tothit;
  int *last_moved, last_totmoved;
  float last_area_no[3], last_area_co[3];
  int64_t vertex_dabs;
  float (*scratch)[3];      /* Jacobi buffer for smooth, indexed by vertex */
  unsigned char *iter_flag; /* smooth: vertex moved in this iteration */
  int *moved_stamp;         /* dab serial at which the vertex was last listed in last_moved */
  int dab_serial;

  /* PBVH_GRIDS (pbvh.c:2516-2561, subdiv_ccg.c): "vertices" are grid elements, element index =
   * grid * grid_size^2 + y * grid_size + x; prims are grids */
  int is_grids, totgrid, grid_size;
  int totface, *face_start, *face_num, *grid_face; /* SubdivCCGFace: start_grid_index, num_grids */
  int totedge, *edge_off, *edge_elems;             /* SubdivCCGAdjacentEdge.boundary_coords as element indices */
  int totcvert, *cvert_off, *cvert_elems;          /* SubdivCCGAdjacentVertex.corner_coords */
  int *grid_edge, *grid_cvert;                     /* coarse edge / vertex at the face corner of each grid */
  int *face_stamp, *edge_stamp, *cvert_stamp, stamp;
  /* stand-ins for the topology refiner's getEdgeVertices / getVertexEdges (or_grids_set_topology) */
  int *edge_verts, *cvert_edge_off, *cvert_edges;
  unsigned char *cvert_boundary;
  int max_neighbors;
};

/* oracle_grids.c */
void or_grids_update_normals(OrPbvh *p, const int *faces, int totface); /* KERNEL_subdiv_ccg_update_normals */
void or_grids_stitch_faces(OrPbvh *p, const int *faces, int totface);   /* KERNEL_subdiv_ccg_average_stitch_faces */
int or_grids_get_updates(OrPbvh *p, int clear, int *r_faces);           /* BKE_pbvh_get_grid_updates */

extern int or_threads;

#endif
