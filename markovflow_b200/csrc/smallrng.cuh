// Counter-based standard-normal draws inside the chain kernels: Philox4x32-10 from the CUDA
// toolkit's device API (curand_kernel.h, header only), keyed so that the draw for (trajectory c, step k,
// component i) does not depend on how the launch is cut into CTAs, tiles or time segments:
//     key = seed,  subsequence = c,  offset = 4 * ceil(D/2) * k 32-bit outputs   (one Philox block = 2 normals)
// The reference draws with TensorFlow's stateless-free tf.random.normal (state_space_model.py:313-316), which has
// no stream a port could pin; mf_philox_normal writes THIS stream out so that a sample can be reproduced
// through sample_from_epsilons.
#pragma once
#include <curand_kernel.h>

#include <cstdint>

namespace mf {

struct ChainRng {
  curandStatePhilox4_32_10_t st;
  __device__ __forceinline__ void init(unsigned long long seed, long long traj, long long step, int d) {
    curand_init(seed, (unsigned long long)traj, 4ull * (unsigned long long)((d + 1) / 2) * (unsigned long long)step, &st);
  }
  // D standard normals of one step (ceil(D/2) Philox blocks, Box-Muller in double precision)
  template <typename T, int D>
  __device__ __forceinline__ void draw(T* e) {
#pragma unroll
    for (int i = 0; i < D; i += 2) {
      const double2 z = curand_normal2_double(&st);
      e[i] = (T)z.x;
      if (i + 1 < D) e[i + 1] = (T)z.y;
    }
  }
};

}  // namespace mf
