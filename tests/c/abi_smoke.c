/* Proves include/ls3d.h is a C header (compiled as C99, -pedantic -Werror) and that a C caller can link libls3d_b200
 * and use the Mesh lifetime calls of the reference boundary (depthprocessing.cpp:1818-1835) without any device.
 * Built and run by tests/test_abi.py::test_header_compiles_and_links_as_c. */
#include <stdio.h>
#include <string.h>
#include "ls3d.h"

int main(void)
{
	Mesh *m;
	Mesh on_stack;
	const char *e;
	if (sizeof(Mesh) != 32 || sizeof(VertexC4ubV3f) != 16 || sizeof(Point3f) != 12 || sizeof(RGB) != 4) {
		printf("layout mismatch: Mesh %u VertexC4ubV3f %u Point3f %u RGB %u\n", (unsigned)sizeof(Mesh), (unsigned)sizeof(VertexC4ubV3f),
			(unsigned)sizeof(Point3f), (unsigned)sizeof(RGB));
		return 1;
	}
	m = createMesh();
	if (!m || m->nVertices != 0 || m->nTriangles != 0 || m->vertices != NULL || m->triangles != NULL) { printf("createMesh: bad initial state\n"); return 2; }
	deleteMesh(m);                 /* nothing to release; the struct itself stays allocated, as in the reference */
	memset(&on_stack, 0, sizeof on_stack);
	deleteMesh(&on_stack);         /* the C# caller passes a by-ref struct (KinectServer.cs:59-60) */
	e = ls3d_last_error();
	if (!e) { printf("ls3d_last_error returned NULL\n"); return 3; }
	printf("ok %s\n", e[0] ? e : "(no error)");
	return 0;
}
