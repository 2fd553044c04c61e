// front_gpu.cu -- device backend for the dense work of the large multifrontal fronts (setup path:
// System::initialize() / recompute_weights() in the reference, System.cpp:138-140, 159-179).
//
// Almost all flops of the sparse Cholesky factorisation sit in the few dense fronts near the root of the
// nested-dissection tree (at 1 M tets: 31 fronts of order 1 900 ... 4 700, 1.5e11 flops; at 8 M tets the root
// front has order 18 500).  The host assembles such a front in a page-locked buffer; here it is factored on the
// B200 in FP64:
//     L11 = chol(F11)                 cusolverDnDpotrf
//     L21 = F21 L11^-T                cublasDtrsm
//     S   = F22 - L21 L21^T           cublasDsyrk          (Schur complement for the parent)
//     X   = inv(L11)                  cublasDtrsm on I
//     T21 = L21 X                     cublasDtrmm
// and T = [X; T21] -- the inverse-multifrontal panel the solve kernels stream -- and S are copied back.
// These are plain library calls on a setup path; the per-iteration hot path (direct_solve.cu) is hand-written.
// cuSOLVER / cuBLAS are resolved with dlopen so that libadmm_b200.so has no link-time dependency on them (a Python
// process may already hold torch's copies; by soname we get whichever is loaded).  If they cannot be loaded the
// factorisation simply stays on the host.
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "direct_factor.h"

namespace admmb {

namespace {

struct Libs {
	bool tried = false, ok = false;
	cublasStatus_t (*blasCreate)(cublasHandle_t *) = nullptr;
	cublasStatus_t (*blasDestroy)(cublasHandle_t) = nullptr;
	cublasStatus_t (*blasSetStream)(cublasHandle_t, cudaStream_t) = nullptr;
	cublasStatus_t (*Dtrsm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int, int, const double *,
	                        const double *, int, double *, int) = nullptr;
	cublasStatus_t (*Dsyrk)(cublasHandle_t, cublasFillMode_t, cublasOperation_t, int, int, const double *, const double *, int, const double *,
	                        double *, int) = nullptr;
	cublasStatus_t (*Dtrmm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int, int, const double *,
	                        const double *, int, const double *, int, double *, int) = nullptr;
	cusolverStatus_t (*solverCreate)(cusolverDnHandle_t *) = nullptr;
	cusolverStatus_t (*solverDestroy)(cusolverDnHandle_t) = nullptr;
	cusolverStatus_t (*solverSetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
	cusolverStatus_t (*potrfBufferSize)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, int *) = nullptr;
	cusolverStatus_t (*potrf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, double *, int, int *) = nullptr;
};
Libs g_libs;

template <class F>
bool sym(void *h, const char *name, F &f) {
	f = reinterpret_cast<F>(dlsym(h, name));
	return f != nullptr;
}

bool load_libs() {
	if (g_libs.tried) return g_libs.ok;
	g_libs.tried = true;
	void *hb = dlopen("libcublas.so.12", RTLD_NOW | RTLD_LOCAL);
	void *hs = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_LOCAL);
	if (!hb || !hs) return false;
	bool ok = true;
	ok &= sym(hb, "cublasCreate_v2", g_libs.blasCreate);
	ok &= sym(hb, "cublasDestroy_v2", g_libs.blasDestroy);
	ok &= sym(hb, "cublasSetStream_v2", g_libs.blasSetStream);
	ok &= sym(hb, "cublasDtrsm_v2", g_libs.Dtrsm);
	ok &= sym(hb, "cublasDsyrk_v2", g_libs.Dsyrk);
	ok &= sym(hb, "cublasDtrmm_v2", g_libs.Dtrmm);
	ok &= sym(hs, "cusolverDnCreate", g_libs.solverCreate);
	ok &= sym(hs, "cusolverDnDestroy", g_libs.solverDestroy);
	ok &= sym(hs, "cusolverDnSetStream", g_libs.solverSetStream);
	ok &= sym(hs, "cusolverDnDpotrf_bufferSize", g_libs.potrfBufferSize);
	ok &= sym(hs, "cusolverDnDpotrf", g_libs.potrf);
	g_libs.ok = ok;
	return ok;
}

// top w x w block of T (ld = m) := identity
__global__ void k_identity_block(int w, long ld, double *T) {
	const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long)w * w) return;
	const int c = (int)(t / w), r = (int)(t - (long)c * w);
	T[r + (long)c * ld] = (r == c) ? 1.0 : 0.0;
}

class DeviceFrontBackend : public FrontBackend {
public:
	DeviceFrontBackend(int device, cudaStream_t stream, int min_front) : device_(device), stream_(stream), min_front_(min_front) {}
	~DeviceFrontBackend() override {
		cudaSetDevice(device_);
		if (blas_) g_libs.blasDestroy(blas_);
		if (solver_) g_libs.solverDestroy(solver_);
		if (h_front_) cudaFreeHost(h_front_);
		if (d_F_) cudaFree(d_F_);
		if (d_T_) cudaFree(d_T_);
		if (d_work_) cudaFree(d_work_);
		if (d_info_) cudaFree(d_info_);
	}
	bool init() { // lazy: the first large front pays for loading the libraries and creating the handles (~0.5 s)
		if (init_tried_) return init_ok_;
		init_tried_ = true;
		init_ok_ = init_once();
		return init_ok_;
	}
	bool init_once() {
		if (!load_libs()) return false;
		if (g_libs.blasCreate(&blas_) != CUBLAS_STATUS_SUCCESS) { blas_ = nullptr; return false; }
		if (g_libs.solverCreate(&solver_) != CUSOLVER_STATUS_SUCCESS) { solver_ = nullptr; return false; }
		g_libs.blasSetStream(blas_, stream_);
		g_libs.solverSetStream(solver_, stream_);
		return cudaMalloc((void **)&d_info_, sizeof(int)) == cudaSuccess;
	}
	int min_front() const override { return min_front_; }
	void reserve(size_t front_doubles, size_t panel_doubles) override {
		if (!init()) return;
		front_buffer(front_doubles);
		grow(d_F_, cap_F_, front_doubles);
		grow(d_T_, cap_T_, panel_doubles);
	}
	double *front_buffer(size_t doubles) override {
		if (!init()) return nullptr;
		if (doubles > h_cap_) {
			if (h_front_) cudaFreeHost(h_front_);
			h_front_ = nullptr;
			h_cap_ = 0;
			if (cudaMallocHost((void **)&h_front_, doubles * sizeof(double)) != cudaSuccess) { cudaGetLastError(); h_front_ = nullptr; return nullptr; }
			h_cap_ = doubles;
		}
		return h_front_;
	}
	int factor_front(int m, int w, const double *Fr, double *T, double *U) override {
		const int r = m - w;
		if (!grow(d_F_, cap_F_, (size_t)m * m) || !grow(d_T_, cap_T_, (size_t)m * w)) return -1;
		int lwork = 0;
		if (g_libs.potrfBufferSize(solver_, CUBLAS_FILL_MODE_LOWER, w, d_F_, m, &lwork) != CUSOLVER_STATUS_SUCCESS) return -1;
		if (!grow(d_work_, cap_work_, (size_t)std::max(lwork, 1))) return -1;
		const double one = 1.0, minus_one = -1.0;
		if (cudaMemcpyAsync(d_F_, Fr, (size_t)m * m * sizeof(double), cudaMemcpyHostToDevice, stream_) != cudaSuccess) return -1;
		if (g_libs.potrf(solver_, CUBLAS_FILL_MODE_LOWER, w, d_F_, m, d_work_, lwork, d_info_) != CUSOLVER_STATUS_SUCCESS) return -1;
		if (r > 0) {
			if (g_libs.Dtrsm(blas_, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, r, w, &one, d_F_, m, d_F_ + w, m) !=
			    CUBLAS_STATUS_SUCCESS) return -1;
			if (U && g_libs.Dsyrk(blas_, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, r, w, &minus_one, d_F_ + w, m, &one, d_F_ + w + (size_t)w * m, m) !=
			             CUBLAS_STATUS_SUCCESS) return -1;
		}
		k_identity_block<<<(unsigned)(((long)w * w + 255) / 256), 256, 0, stream_>>>(w, m, d_T_);
		if (g_libs.Dtrsm(blas_, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, w, w, &one, d_F_, m, d_T_, m) !=
		    CUBLAS_STATUS_SUCCESS) return -1;
		if (r > 0 && g_libs.Dtrmm(blas_, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, r, w, &one, d_T_, m, d_F_ + w, m,
		                          d_T_ + w, m) != CUBLAS_STATUS_SUCCESS) return -1;
		int info = 0;
		cudaMemcpyAsync(&info, d_info_, sizeof(int), cudaMemcpyDeviceToHost, stream_);
		cudaMemcpyAsync(T, d_T_, (size_t)m * w * sizeof(double), cudaMemcpyDeviceToHost, stream_);
		if (U && r > 0)
			cudaMemcpy2DAsync(U, (size_t)r * sizeof(double), d_F_ + w + (size_t)w * m, (size_t)m * sizeof(double), (size_t)r * sizeof(double), r,
			                  cudaMemcpyDeviceToHost, stream_);
		if (cudaStreamSynchronize(stream_) != cudaSuccess) { cudaGetLastError(); return -1; }
		return info != 0 ? 1 : 0;
	}

private:
	bool grow(double *&p, size_t &cap, size_t need) {
		if (need <= cap) return true;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		if (cudaMalloc((void **)&p, need * sizeof(double)) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
		cap = need;
		return true;
	}
	int device_;
	cudaStream_t stream_;
	int min_front_;
	cublasHandle_t blas_ = nullptr;
	cusolverDnHandle_t solver_ = nullptr;
	double *h_front_ = nullptr, *d_F_ = nullptr, *d_T_ = nullptr, *d_work_ = nullptr;
	size_t h_cap_ = 0, cap_F_ = 0, cap_T_ = 0, cap_work_ = 0;
	int *d_info_ = nullptr;
	bool init_tried_ = false, init_ok_ = false;
};

} // namespace

// nullptr when the libraries are unavailable or ADMMB_HOST_FACTOR is set: the factorisation then runs on the host cores.
FrontBackend *make_device_front_backend(admmb_ctx *ctx) {
	if (const char *e = getenv("ADMMB_HOST_FACTOR"))
		if (e[0] && e[0] != '0') return nullptr;
	int min_front = 1536; // below this the host's tree-parallel path is as fast as a device round trip
	if (const char *e = getenv("ADMMB_GPU_FRONT_MIN")) { const int v = atoi(e); if (v >= 16) min_front = v; }
	return new DeviceFrontBackend(ctx->device, ctx->stream, min_front);
}

} // namespace admmb
