// kdbuild.cpp -- host-side producer of the reference-shaped kd-tree.
//
// Replaces KdTree::new -> construct -> median (ekzhang/rpt src/kdtree.rs:108-119,
// 235-355).  The GPU traversal consumes exactly the tree the reference would build
// (same medians, same inclusive two-sided partition, same 0.85 score cut, same
// `< 16` leaf rule, leaf refs in ascending triangle order), so closest-hit parity
// holds triangle for triangle.
//
// Not a transcription: the reference re-derives every bounding box at every node and
// fully sorts three 2n-element endpoint arrays to read one median each
// (O(n log^2 n)).  Here the boxes are computed once, each median is taken with two
// selection passes (nth_element + max of the lower half -- the same two order
// statistics a sort would expose), and disjoint subtrees are built in parallel with
// OpenMP tasks.  Nodes are emitted in depth-first pre-order (left child = node + 1).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/rpt_b200.h"

namespace {

struct Box {
    double lo[3], hi[3];
};

struct Node {
    int kind = 3;
    double split = 0.0;
    std::unique_ptr<Node> left, right;
    std::vector<uint32_t> refs;
};

// median of the multiset {lo[i][axis], hi[i][axis]} exactly as median(sorted) would give it
double endpoint_median(const std::vector<Box>& boxes, const std::vector<uint32_t>& idx, int axis,
                       std::vector<double>& scratch) {
    const size_t n2 = idx.size() * 2;
    scratch.resize(n2);
    for (size_t i = 0; i < idx.size(); i++) {
        scratch[2 * i] = boxes[idx[i]].lo[axis];
        scratch[2 * i + 1] = boxes[idx[i]].hi[axis];
    }
    const size_t mid = n2 / 2;  // n2 is even: (s[mid] + s[mid-1]) / 2
    std::nth_element(scratch.begin(), scratch.begin() + mid, scratch.end());
    const double upper = scratch[mid];
    const double lower = *std::max_element(scratch.begin(), scratch.begin() + mid);
    return (upper + lower) / 2.0;
}

std::unique_ptr<Node> build(const std::vector<Box>& boxes, std::vector<uint32_t> idx) {
    auto node = std::make_unique<Node>();
    const size_t n = idx.size();
    if (n < 16) {
        node->refs = std::move(idx);
        return node;
    }
    std::vector<double> scratch;
    double med[3];
    size_t score[3];
    double blo[3] = {1.0 / 0.0, 1.0 / 0.0, 1.0 / 0.0}, bhi[3] = {-1.0 / 0.0, -1.0 / 0.0, -1.0 / 0.0};
    for (int a = 0; a < 3; a++) {
        med[a] = endpoint_median(boxes, idx, a, scratch);
        size_t l = 0, r = 0;
        for (uint32_t i : idx) {
            const Box& b = boxes[i];
            l += b.lo[a] <= med[a];
            r += b.hi[a] >= med[a];
            blo[a] = std::min(blo[a], b.lo[a]);
            bhi[a] = std::max(bhi[a], b.hi[a]);
        }
        score[a] = std::max(l, r);
    }
    const size_t threshold = (size_t)((double)n * 0.85);
    if (std::min(std::min(score[0], score[1]), score[2]) >= threshold) {
        node->refs = std::move(idx);
        return node;
    }
    int dir = -1;
    const double ex = bhi[0] - blo[0], ey = bhi[1] - blo[1], ez = bhi[2] - blo[2];
    if (ex > ey && ex > ez) {
        if (score[0] < threshold) dir = 0;
    } else if (ey > ez) {
        if (score[1] < threshold) dir = 1;
    } else if (score[2] < threshold) {
        dir = 2;
    }
    if (dir == -1) {
        if (score[0] < score[1] && score[0] < score[2]) dir = 0;
        else if (score[1] < score[2]) dir = 1;
        else dir = 2;
    }
    std::vector<uint32_t> l, r;
    l.reserve(n);
    r.reserve(n);
    for (uint32_t i : idx) {
        if (boxes[i].lo[dir] <= med[dir]) l.push_back(i);
        if (boxes[i].hi[dir] >= med[dir]) r.push_back(i);
    }
    idx.clear();
    idx.shrink_to_fit();
    node->kind = dir;
    node->split = med[dir];
    Node* np = node.get();
    if (n > 20000) {
#pragma omp task shared(boxes) firstprivate(np) untied
        np->left = build(boxes, std::move(l));
#pragma omp task shared(boxes) firstprivate(np) untied
        np->right = build(boxes, std::move(r));
#pragma omp taskwait
    } else {
        np->left = build(boxes, std::move(l));
        np->right = build(boxes, std::move(r));
    }
    return node;
}

void emit(const Node& n, std::vector<rptb_kdnode>& nodes, std::vector<uint32_t>& refs, uint32_t depth,
          uint32_t& max_depth, uint32_t& max_leaf) {
    const size_t me = nodes.size();
    nodes.emplace_back();
    std::memset(&nodes[me], 0, sizeof(rptb_kdnode));
    max_depth = std::max(max_depth, depth);
    if (n.kind == 3) {
        nodes[me].kind = 3;
        nodes[me].first_ref = (uint32_t)refs.size();
        nodes[me].num_refs = (uint32_t)n.refs.size();
        max_leaf = std::max(max_leaf, (uint32_t)n.refs.size());
        refs.insert(refs.end(), n.refs.begin(), n.refs.end());
        return;
    }
    nodes[me].kind = (uint32_t)n.kind;
    nodes[me].split = n.split;
    nodes[me].left = (uint32_t)nodes.size();
    emit(*n.left, nodes, refs, depth + 1, max_depth, max_leaf);
    nodes[me].right = (uint32_t)nodes.size();
    emit(*n.right, nodes, refs, depth + 1, max_depth, max_leaf);
}

}  // namespace

namespace rptb {

static int build_from_boxes(std::vector<Box>& boxes, std::vector<rptb_kdnode>& nodes, std::vector<uint32_t>& refs,
                            uint32_t& depth, uint32_t& max_leaf) {
    std::vector<uint32_t> idx(boxes.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (uint32_t)i;
    std::unique_ptr<Node> root;
#pragma omp parallel
#pragma omp single
    root = build(boxes, std::move(idx));
    nodes.clear();
    refs.clear();
    depth = 0;
    max_leaf = 0;
    emit(*root, nodes, refs, 0, depth, max_leaf);
    return 0;
}

int build_kdtree_host(const double* tris, uint64_t ntris, std::vector<rptb_kdnode>& nodes, std::vector<uint32_t>& refs,
                      uint32_t& depth, uint32_t& max_leaf) {
    std::vector<Box> boxes(ntris);
    for (uint64_t i = 0; i < ntris; i++) {  // Triangle::bounding_box, src/shape/mesh.rs:39-46
        const double* t = tris + 18 * i;
        for (int a = 0; a < 3; a++) {
            boxes[i].lo[a] = std::fmin(std::fmin(t[a], t[3 + a]), t[6 + a]);
            boxes[i].hi[a] = std::fmax(std::fmax(t[a], t[3 + a]), t[6 + a]);
        }
    }
    return build_from_boxes(boxes, nodes, refs, depth, max_leaf);
}

// KdTree<T: Bounded>::new over caller-supplied bounding boxes (p_min, p_max per object).
int build_kdtree_boxes_host(const double* in, uint64_t nboxes, std::vector<rptb_kdnode>& nodes, std::vector<uint32_t>& refs,
                            uint32_t& depth, uint32_t& max_leaf) {
    std::vector<Box> boxes(nboxes);
    for (uint64_t i = 0; i < nboxes; i++)
        for (int a = 0; a < 3; a++) {
            boxes[i].lo[a] = in[6 * i + a];
            boxes[i].hi[a] = in[6 * i + 3 + a];
        }
    return build_from_boxes(boxes, nodes, refs, depth, max_leaf);
}

}  // namespace rptb
