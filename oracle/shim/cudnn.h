// TEST INFRASTRUCTURE (oracle). Stub of the cuDNN surface named by the reference's
// src/cudnn_traits.hpp and src/cudnn_kernel_pool.hpp.  peak_finder_t's constructor always creates a
// pooling descriptor (src/post-process.h:151) even when use_gpu=false, so the symbols must exist; the
// use_gpu=false path never runs cudnnPoolingForward.  Every call succeeds and does nothing.
#pragma once
typedef enum { CUDNN_STATUS_SUCCESS = 0 } cudnnStatus_t;
typedef enum { CUDNN_TENSOR_NCHW = 0, CUDNN_TENSOR_NHWC = 1 } cudnnTensorFormat_t;
typedef enum { CUDNN_DATA_FLOAT = 0, CUDNN_DATA_DOUBLE = 1 } cudnnDataType_t;
typedef enum { CUDNN_CONVOLUTION = 0, CUDNN_CROSS_CORRELATION = 1 } cudnnConvolutionMode_t;
typedef enum { CUDNN_POOLING_MAX = 0 } cudnnPoolingMode_t;
typedef enum { CUDNN_NOT_PROPAGATE_NAN = 0 } cudnnNanPropagation_t;

struct cudnnContext { int unused; };
struct cudnnTensorStruct { int n, c, h, w; };
struct cudnnFilterStruct { int unused; };
struct cudnnConvolutionStruct { int unused; };
struct cudnnPoolingStruct { int unused; };
typedef cudnnContext *cudnnHandle_t;
typedef cudnnTensorStruct *cudnnTensorDescriptor_t;
typedef cudnnFilterStruct *cudnnFilterDescriptor_t;
typedef cudnnConvolutionStruct *cudnnConvolutionDescriptor_t;
typedef cudnnPoolingStruct *cudnnPoolingDescriptor_t;

inline cudnnStatus_t cudnnCreate(cudnnHandle_t *h) { *h = new cudnnContext(); return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroy(cudnnHandle_t h) { delete h; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnCreateTensorDescriptor(cudnnTensorDescriptor_t *d) { *d = new cudnnTensorStruct(); return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyTensorDescriptor(cudnnTensorDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnSetTensor4dDescriptor(cudnnTensorDescriptor_t d, cudnnTensorFormat_t, cudnnDataType_t, int n, int c, int h, int w) { d->n = n; d->c = c; d->h = h; d->w = w; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnGetTensor4dDescriptor(cudnnTensorDescriptor_t d, cudnnDataType_t *t, int *n, int *c, int *h, int *w, int *ns, int *cs, int *hs, int *ws) { *t = CUDNN_DATA_FLOAT; *n = d->n; *c = d->c; *h = d->h; *w = d->w; *ws = 1; *hs = d->w; *cs = d->h * d->w; *ns = d->c * d->h * d->w; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyFilterDescriptor(cudnnFilterDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyConvolutionDescriptor(cudnnConvolutionDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnCreatePoolingDescriptor(cudnnPoolingDescriptor_t *d) { *d = new cudnnPoolingStruct(); return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyPoolingDescriptor(cudnnPoolingDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnSetPoolingNdDescriptor(cudnnPoolingDescriptor_t, cudnnPoolingMode_t, cudnnNanPropagation_t, int, const int *, const int *, const int *) { return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnGetPooling2dForwardOutputDim(cudnnPoolingDescriptor_t, cudnnTensorDescriptor_t x, int *n, int *c, int *h, int *w) { *n = x->n; *c = x->c; *h = x->h; *w = x->w; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnPoolingForward(cudnnHandle_t, cudnnPoolingDescriptor_t, const void *, cudnnTensorDescriptor_t, const void *, const void *, cudnnTensorDescriptor_t, void *) { return CUDNN_STATUS_SUCCESS; }
