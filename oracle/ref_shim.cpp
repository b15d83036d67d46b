// ORACLE (test infrastructure).  C-callable shim around the reference's OWN header-only C++
// evaluator, compiled from the sources where they lie under /root/reference (never copied):
//   evaluator/backend/cpp/include/evaluate.h:45-64  (cpp_evaluate_matrix, thread pool, partial_sort_copy)
//   evaluator/backend/cpp/include/metric.h:17-114   (precision / recall / ap / ndcg / mrr)
// Built by oracle/Makefile into oracle/_ref/libref_eval.so (git-ignored, travels to the GPU box).
#include "evaluate.h"

extern "C" void ref_evaluate_matrix(float* scores, int n_users, int rating_len, const int* truth_ptr,
                                    const int* truth_items, const int* metric_ids, int n_metrics, int top_k,
                                    int threads, float* results) {
    std::vector<std::unordered_set<int>> truth(n_users);
    for (int u = 0; u < n_users; ++u)
        for (int e = truth_ptr[u]; e < truth_ptr[u + 1]; ++e) truth[u].insert(truth_items[e]);
    std::vector<int> metric(metric_ids, metric_ids + n_metrics);
    cpp_evaluate_matrix(scores, rating_len, truth, metric, top_k, threads, results);
}
