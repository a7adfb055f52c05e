/* TEST INFRASTRUCTURE ONLY.  Stand-ins for the two C symbols the reference's Cython
 * inject_particles binds (cdef extern draw_velocities_c / sgenrand, infMagSim_cython.py:224-225),
 * used by the SECOND build of oracle/build_inject_ref.py (inject_ref_given): the velocities are
 * handed in by the test instead of being drawn, so that the reference's injection body can be
 * compared bit for bit with the CUDA path on the CUDA path's own (Philox) velocities.  Nothing
 * here restates reference arithmetic. */
#include <string.h>

static const float *g_velocities = 0;
static int g_count = 0;

void inject_shim_set_velocities(const float *v, int n) {
    g_velocities = v;
    g_count = n;
}

void draw_velocities_c(int n, float v_th, float v_d, float *v_array) {
    (void)v_th;
    (void)v_d;
    if (n > g_count) n = g_count;
    if (n > 0) memcpy(v_array, g_velocities, (size_t)n * sizeof(float));
}

void sgenrand(unsigned long seed) { (void)seed; }
