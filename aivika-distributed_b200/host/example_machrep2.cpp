// C++ twin of tests/MachRep2DistributedReproducible.hs `master`: time server with quorum 3, one master and two slave
// LPs, default DIOParams (no time horizon), stop time 1000.  Build (needs a B200 to run):
//   g++ -std=c++17 example_machrep2.cpp -L../csrc -ltwb200 -Wl,-rpath,'$ORIGIN/../csrc' -o example_machrep2
#include <cstdio>
#include "dio.hpp"

using namespace aivika::distributed;

int main()
{
    try {
        TimeServer ts = timeServer(3);
        DIOParams ps = defaultDIOParams;
        ps.queueCapacity = 64; ps.stepsPerLaunch = 32;
        auto [dio, st] = runDIO(MachRep2{1}, ps, ts, Specs{0.0, 1000.0, 1.0});
        auto r = dio->machRepResults();
        std::printf("The result is (%.17g,%.17g)\n", r[0].up_time_prop, r[0].immed_prop);     // "about 0.6 and 0.67"
        std::printf("committed %llu, rolled back %llu, GVT %.6f\n", (unsigned long long)st.committed,
                    (unsigned long long)st.rolled_back, *ts.globalTime());
    } catch (const SimulationAbort &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
