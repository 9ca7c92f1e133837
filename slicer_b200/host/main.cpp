// SLICER_b200 — same command line as the reference (README.md:80-84):   SLICER_b200 InputParams.ini
// Optional switches stand in for the reference's compile-time options:
//   --gpus 0,1,..      GPUs of this box that share the sub-files (default 0)        [MPI ranks, slicer-v2.cpp:28-30]
//   --replication      CMake USE_REPLICATION   (-DReplicationOnPerpendicularPlane, CMakeLists.txt:20-22)
//   --fixed-vertex     CMake USE_FIXED_PLC_VERTEX (-DFixedPLCVertex, CMakeLists.txt:23-26)
//   --ngp              `#define DO_NGP true`   (densitymaps.h:22)
#include "slicer_host.h"

#include <cstring>
#include <iostream>

int main(int argc, char **argv)
{
  slicer::RunOptions opt;
  std::string ini;
  for (int i = 1; i < argc; i++)
  {
    if (!strcmp(argv[i], "--gpus") && i + 1 < argc)
    {
      opt.devices.clear();
      for (char *t = strtok(argv[++i], ","); t; t = strtok(nullptr, ","))
        opt.devices.push_back(atoi(t));
    }
    else if (!strcmp(argv[i], "--replication"))
      opt.replication = true;
    else if (!strcmp(argv[i], "--fixed-vertex"))
      opt.fixed_vertex = true;
    else if (!strcmp(argv[i], "--ngp"))
      opt.mas = SLICER_MAS_NGP;
    else if (!strcmp(argv[i], "--quiet"))
      opt.quiet = true;
    else
      ini = argv[i];
  }
  if (ini.empty())
  {
    std::cerr << "usage: SLICER_b200 [--gpus 0,1,..] [--replication] [--fixed-vertex] [--ngp] InputParams.ini" << std::endl;
    return 1;
  }
  return slicer::runLightCone(ini, opt) ? 255 : 0; // the reference ends in MPI_Abort(-1) on every error
}
