// ref_sampler_wrapper.cpp — extern "C" shim over the UNMODIFIED reference sampler so that tests / bench can call
// it through ctypes.  The reference source is compiled where it lies (REF_SAMPLER_CPP is passed by the Makefile as
// an absolute path under /root/reference); nothing from it is copied into this repository.
// TEST INFRASTRUCTURE ONLY (oracle/): never linked into libnncf_b200.so.
#include REF_SAMPLER_CPP

extern "C" {
void* ref_sampler_create(const double* dist, int dist_size, double power, unsigned long long seed) {
  return new nodesampler::NodeSampler(dist, dist_size, power, seed);
}
int ref_sampler_sample(void* h) { return static_cast<nodesampler::NodeSampler*>(h)->sample(); }
void ref_sampler_sample_batch(void* h, int n, int* out) {
  static_cast<nodesampler::NodeSampler*>(h)->sample_batch(n, out);
}
}
