// CPU-only check of the rig XML reader in vicalib_b200/host/vicalibrator.h (no CUDA call is made: the
// ViCalibrator class itself is not instantiated).  Prints what it parsed, one camera per line.
#include <cstdio>

#include "../../vicalib_b200/host/vicalibrator.h"

using namespace visual_inertial_calibration;

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s rig.xml\n", argv[0]); return 2; }
  try {
    const auto rig = ReadXmlRig(argv[1]);
    for (const auto& c : rig) {
      std::printf("%s %d %d %d params", c->Type().c_str(), c->Index(), c->Width(), c->Height());
      for (double v : c->GetParams()) std::printf(" %.17g", v);
      std::printf(" rdf");
      for (int k = 0; k < 9; ++k) std::printf(" %.17g", c->RDF()[k]);
      double M[12];
      c->Pose().matrix3x4(M);
      std::printf(" T_wc");
      for (int k = 0; k < 12; ++k) std::printf(" %.17g", M[k]);
      std::printf("\n");
    }
  } catch (const std::exception& e) {
    std::printf("error %s\n", e.what());
    return 1;
  }
  return 0;
}
