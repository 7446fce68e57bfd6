/* Empty stand-in for MSVC's <intrin.h>, which the reference includes beside
 * <immintrin.h> (och_h_octree.h:7, och_octree.cpp:7).  g++ gets everything it
 * needs from <immintrin.h>.  Test infrastructure only. */
#pragma once
