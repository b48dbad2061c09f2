// Flat C entry points over the reference's C++-only filter() (test infrastructure only); compiled against
// the reference's own headers and linked with its own filter.o into oracle/_ref/libls3d_ref_filter.so.
#include "filter.h"

std::vector<KNNeighborsResult> KNNeighbors(PointCloud &cloud, kdTree &tree, int k);   // filter.cpp:19

extern "C" {

// filter (src/LiveScanClient/filter.cpp:36-81). verts/colors are compacted in place exactly as the reference
// leaves its vectors; old_to_new[i] is the returned map's entry for i (or -2 where the map has no entry,
// i.e. on the k<=0 || maxDist<=0 early return).  Returns the surviving count.
int ref_filter(float *verts, unsigned char *colors, int n, int k, float maxDist, int *old_to_new)
{
	std::vector<Point3f> v((Point3f*)verts, (Point3f*)verts + n);
	std::vector<RGB> c((RGB*)colors, (RGB*)colors + n);
	std::unordered_map<int, int> m = filter(v, c, k, maxDist);
	for (int i = 0; i < n; i++) { auto it = m.find(i); old_to_new[i] = (it == m.end()) ? -2 : it->second; }
	if (!v.empty()) { memcpy(verts, v.data(), v.size() * sizeof(Point3f)); memcpy(colors, c.data(), c.size() * sizeof(RGB)); }
	return (int)v.size();
}

// KNNeighbors (filter.cpp:19-34): squared distance to the k-th nearest neighbour (self included) of every point.
void ref_knn_kdist(float *verts, int n, int k, float *kdist)
{
	PointCloud cloud;
	cloud.pts = std::vector<Point3f>((Point3f*)verts, (Point3f*)verts + n);
	kdTree tree(3, cloud);
	tree.buildIndex();
	std::vector<KNNeighborsResult> r = KNNeighbors(cloud, tree, k);
	for (int i = 0; i < n; i++) kdist[i] = r[i].kDistance;
}

}  // extern "C"
