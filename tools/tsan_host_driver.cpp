// TSan driver: threaded fixture builder + flatten + edits, host code only
#include "ort_b200.h"
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <vector>
#include <random>
struct ort_ctx;
int ort_fail(ort_ctx*, int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }
extern "C" int ort_upload_full(ort_ctx*, const uint32_t*, size_t, uint32_t) { return ORT_ERR_NOT_ATTACHED; }
extern "C" int ort_upload_delta(ort_ctx*, const uint32_t*, const uint32_t*, size_t, uint32_t) { return ORT_ERR_NOT_ATTACHED; }
extern "C" int ort_upload_pool(ort_ctx*, const uint32_t*, size_t) { return ORT_ERR_NOT_ATTACHED; }
int main(int argc, char** argv)
{
	const int depth = argc > 1 ? atoi(argv[1]) : 9, log2cap = argc > 2 ? atoi(argv[2]) : 21, nthreads = 8;
	const int dim = 1 << depth;
	std::vector<uint16_t> h(static_cast<size_t>(dim) * dim);
	ort_fixture_heightmap(depth, h.data(), nthreads);
	std::vector<uint8_t> grass(static_cast<size_t>(dim) * dim);
	std::mt19937 rng(1);
	for (auto& g : grass) g = rng() & 1;
	for (int tunnels = 0; tunnels < 2; ++tunnels)
	{
		ort_tree* t = nullptr;
		if (ort_tree_create(&t, log2cap, depth)) return 1;
		if (ort_fixture_build_terrain(t, h.data(), grass.data(), tunnels, nthreads)) return 2;
		const uint32_t* nodes = nullptr; uint32_t root = 0;
		const size_t n = ort_tree_flatten(t, &nodes, &root, nullptr);
		printf("depth %d tunnels %d: %zu nodes, root %u, fill %u\n", depth, tunnels, n, root, ort_tree_get_fillcnt(t));
		ort_tree_fill_box(t, 10, 10, 10, 50, 50, 50, 1);
		ort_tree_set_box(t, 100, 100, 60, 20, 0);
		const size_t n2 = ort_tree_flatten(t, &nodes, &root, nullptr);
		printf("  after edits: %zu nodes\n", n2);
		ort_tree_destroy(t);
	}
	return 0;
}
