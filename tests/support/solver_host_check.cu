// Host-side run of the device solver's stack machine (csrc/rz_solver.cuh is host/device code): reads lines
// "own enemy exactly" (hex hex int), prints "move score".  argv[1] = deadline step in fake-clock ticks (0: run each request
// to completion in one call; n > 0: suspend/resume roughly every 16 node steps, as the engine's time-sliced solver step does).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rz_solver.cuh"
using namespace rz;
using namespace rz::solver;
int main(int argc, char** argv) {
    const long long slice = argc > 1 ? atoll(argv[1]) : 0;
    std::vector<u64> table(kTtEntries * kTtWordsPerEntry, 0);
    const TT tt{table.data()};
    unsigned long long own, enemy;
    int exactly;
    SolveCtx* c = new SolveCtx;
    long long resumes = 0;
    while (scanf("%llx %llx %d", &own, &enemy, &exactly) == 3) {
        ctx_init(c, own, enemy, exactly);
        while (!solve_advance(c, tt, slice ? global_ns() + slice : 0)) ++resumes;
        printf("%d %d\n", (int)c->move, c->move < 0 ? 0 : (int)c->score);
    }
    fprintf(stderr, "resumes %lld\n", resumes);
    delete c;
    return 0;
}
