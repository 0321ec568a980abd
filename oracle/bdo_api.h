// Output block shared by the oracle (oracle/bd_oracle.cpp) and the host simulation of the device
// logic (tests/hostsim/hostsim.cpp). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <stdint.h>
#include "../include/bdk.h"

extern "C" {
struct bdo_output {
    bdk_summary_t summary;
    uint64_t n_sv;
    bdk_sv* sv;
    int32_t* lib_count;     // [n_sv][nlib]
    uint32_t* cn_count;     // [n_sv][nkey]
    float* copy_number;     // [n_sv][nkey]
    int32_t nkey;
    uint64_t n_regions;
    bdk_region* regions;
    uint8_t* region_alive;  // [n_regions] region still exists at exit
    uint64_t n_areads;
    bdk_aread* areads;
    int32_t* aread_region;
    int32_t* sv_of_read;    // [n_areads] order of the emitted SV whose support list holds the read, else -1
    uint8_t* rec_class;     // [n] pass-2 flag per input record, 255 = filtered out
    uint64_t n_support;     // total support entries
    uint64_t* support_off;  // [n_sv+1]
    uint32_t* support;      // record indices, support_reads order
    int32_t n_flush;
};

}
