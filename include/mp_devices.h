/*
 * mp_devices.h -- device table, streams, target-device state, work submission
 * and the per-device worker pools (C ABI).
 *
 * mpdev_*: the twenty entry points of src/include/millipyde_devices.h:21-79
 * (src/millipyde_devices.cpp).  mpwrk_*: the eight of
 * src/include/millipyde_workers.h:40-62 (src/millipyde_workers.cpp); the pool
 * struct is opaque here (the reference exposes its pthread fields but nothing
 * outside millipyde_workers.cpp touches them).
 *
 * B200 differences behind the same calls: every stream (index 0 included) is a
 * non-blocking stream; a device's context, streams and pool are set up by the
 * first call that uses it (a one-GPU job on an 8-GPU box owns one context);
 * peer access and pool access are granted per ordered pair by the first
 * hand-off that needs them (mpdev_can_use_peer answers "yes" by making it so);
 * no cudaDeviceReset during the P2P probe.
 */
#ifndef MP_B200_DEVICES_H
#define MP_B200_DEVICES_H
#include "mp_abi.h"

/* Streams per device, index 0 is the default op stream (the reference's NULL
 * stream slot), 1..4 serve Pipeline/Generator work. millipyde_devices.h:11-13 */
#define DEVICE_STREAM_COUNT 5
#define THREADS_PER_DEVICE ((DEVICE_STREAM_COUNT)-1)

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*MPWorkItem)(void *arg);
typedef struct work_node MPWorkNode;
typedef struct work_pool MPDeviceWorkPool;

MPStatus mpdev_initialize(void);
void mpdev_teardown(void);
MPBool mpdev_peer_to_peer_supported(void);
MPBool mpdev_can_use_peer(int device, int peer_device);
int mpdev_get_device_count(void);
MPBool mpdev_is_valid_device(int device_id);
void *mpdev_get_stream(int device_id, int stream);
void mpdev_submit_work(int device_id, MPWorkItem work, void *arg);
void mpdev_hard_synchronize(int device_id);
void mpdev_hard_synchronize_all(void);
void mpdev_synchronize(void);
void mpdev_synchronize_all(void);
void mpdev_reset(int device_id);
void mpdev_set_device(int device_id);
void mpdev_stream_synchronize(int device_id, int stream_id);
int mpdev_get_target_device(void);
int mpdev_get_alternative_device(int device_id);
int mpdev_get_next_device(int device_id);
void mpdev_set_target_device(int device_id);
int mpdev_get_recommended_device(void);

MPStatus mpwrk_create_work_node(MPWorkNode **result, MPWorkItem work, void *arg);
void mpwrk_destroy_work_node(MPWorkNode *node);
MPWorkNode *mpwrk_work_queue_pop(MPDeviceWorkPool *pool);
MPStatus mpwrk_work_queue_push(MPDeviceWorkPool *pool, MPWorkItem work, void *arg);
void mpwrk_work_wait(MPDeviceWorkPool *pool);
void *mpwrk_process_work(void *arg);
MPStatus mpwrk_create_work_pool(MPDeviceWorkPool **result, int num_threads);
MPStatus mpwrk_destroy_work_pool(MPDeviceWorkPool *pool);

/* ---- new: device-side timing and memory queries for bench.py ------------ */
typedef struct mp_event MPEvent;
MPEvent *mpdev_event_create(int device_id);
void mpdev_event_destroy(MPEvent *ev);
void mpdev_event_record(MPEvent *ev, void *stream);
float mpdev_event_elapsed_ms(MPEvent *start, MPEvent *stop); /* syncs on stop */
MPStatus mpdev_mem_info(int device_id, size_t *free_bytes, size_t *total_bytes);
int mpdev_sm_count(int device_id);
/* PCI bus id of a device ("0000:1b:00.0"): how a monitoring tool (NVML) finds the same GPU */
MPStatus mpdev_pci_bus_id(int device_id, char *buf, int len);
/* write > L2-size bytes so the next timed launch starts cold */
void mpdev_flush_l2(int device_id, void *stream);
/* Hand the unused part of every device pool back to the driver (the pools otherwise keep every
 * freed block: release threshold = never).  For a process that stays alive but is done with the GPU. */
void mpdev_trim_pools(void);
/* number of kernels this library launched since load (all threads) */
unsigned long long mpdev_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MP_B200_DEVICES_H */
