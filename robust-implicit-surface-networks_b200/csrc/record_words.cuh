// Iso record format (shared by the kernels that write and read records).
#pragma once
#include <cstdint>

namespace rin {

// ---------------------------------------------------------------------------------------------
// Iso record: the part of a per-tet complex that the mesh extraction consumes
// (/root/reference/src/extract_mesh.cpp:60-261 reads only iso vertices and iso faces).
// A sequence of 32-bit words (records are 4-byte aligned so that one load fetches one entry):
//   word 0                 n_iso_verts | n_iso_faces << 8 | n_face_vertex_entries << 16
//   n_iso_verts words      local vertex id | plane0 << 8 | plane1 << 16 | plane2 << 24  (ascending)
//   per iso face           local face id (16) | supporting plane << 16 | n << 24 | boundary << 31
//                          followed by ceil(n/4) words of iso-vertex ranks, one byte each
//   trailing word          number of faces of the WHOLE complex (read only by the cell-grouping maps,
//                          src/extract_mesh.cpp:268-566; the extraction kernels stop before it)
// boundary: the face lies on the tet boundary (negative_cell == None).
// Plane ids: 0..3 simplex faces, 4+j the j-th active function of the tet.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t rec_face_words(int n)
{
    return 1u + uint32_t(n + 3) / 4u;
}

} // namespace rin
