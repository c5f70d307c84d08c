// TEST INFRASTRUCTURE ONLY (oracle/): builds the UNMODIFIED reference engine for
// differential testing and for bench.py's cpu_baseline / --impl reference legs.
//
// The reference's GridWorld.cc does not compile as shipped: line 333 of
// examples/battle_model/src/gridworld/GridWorld.cc is a lone token `p` where the
// declarations used at GridWorld.cc:392-394,400 (range, view_*_offset,
// view_left_top_*, view_right_bottom_*, channel_trans) were deleted.
//
// No reference source is copied or patched on disk.  This translation unit
// (1) pre-includes every header GridWorld.cc pulls in (their include guards then
//     make the later #include a no-op, so the macro below cannot touch them),
// (2) defines the stray token `p` as the missing declarations -- each name and
//     type is forced by the surviving call site (GridWorld.cc:392-394,400), by
//     Map::extract_view's parameter list (Map.h:55-58), AgentType.h:35,
//     Range.h:57-60 and make_channel_trans (GridWorld.cc:981-997),
// (3) #includes GridWorld.cc from where it lies (REF_SRC is passed by the Makefile).
//
// `p` occurs as a standalone token exactly once in GridWorld.cc (checked by
// oracle/Makefile before compiling).

#include <iostream>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <cassert>
#include <stdexcept>

#include "gridworld/GridWorld.h"

#define p                                                                              \
    const Range *range = type.view_range;                                              \
    int view_x_offset = type.view_x_offset, view_y_offset = type.view_y_offset;        \
    int view_left_top_x, view_left_top_y, view_right_bottom_x, view_right_bottom_y;    \
    range->get_range_rela_offset(view_left_top_x, view_left_top_y,                     \
                                 view_right_bottom_x, view_right_bottom_y);            \
    std::vector<int> channel_trans =                                                   \
        make_channel_trans(group, group2channel(0), type.n_channel, n_group);

#include "gridworld/GridWorld.cc"

#undef p
