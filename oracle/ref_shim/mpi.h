// TEST INFRASTRUCTURE ONLY (oracle): MPI is not installed here; the reference's call sites only pass the communicator through.
#pragma once
typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
