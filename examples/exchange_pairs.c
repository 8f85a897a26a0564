#define _POSIX_C_SOURCE 200809L
/* Cross-image matching between processes through the library's descriptor exchange (include/vksift_b200_ext.h,
 * vksiftx_exchange*), from plain C and without MPI or torch: the parent forks NPROC workers BEFORE any CUDA call, the workers
 * pass their 64-byte IPC handles and two barriers through pipes.  Every worker detects features on its own synthetic image,
 * then ONE call pushes its descriptors to every peer, waits for theirs and matches against all of them in place.
 * Here the workers share GPU 0 (set gpu_device_index = rank to give each its own GPU on a multi-GPU node).
 *   gcc examples/exchange_pairs.c -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -lm -o exchange_pairs
 * Prints, per rank and peer, how many of its features have their nearest neighbour in the peer's image at a ratio < 0.75, and
 * checks the record invariants (idx_a = row, indices inside the peer's count, dist1 <= dist2). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vksift_b200_ext.h>
#include <vulkansift/vulkansift.h>

#define NPROC 3
#define SLOT_ROWS 2048
#define HANDLE_BYTES 64

static void blobs(uint8_t *img, int w, int h, int shift)
{
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
    {
      float v = 0.5f;
      for (int k = 0; k < 40; k++)
      {
        const float cx = (float)((k * 97 + shift) % w), cy = (float)((k * 57) % h), s = 3.f + (float)(k % 5);
        const float d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        v += ((k & 1) ? 0.4f : -0.4f) * expf(-d2 / (2.f * s * s));
      }
      v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
      img[y * w + x] = (uint8_t)(255.f * v + 0.5f);
    }
}

static int read_all(int fd, void *buf, size_t n)
{
  size_t got = 0;
  while (got < n)
  {
    const ssize_t r = read(fd, (char *)buf + got, n - got);
    if (r <= 0)
      return 0;
    got += (size_t)r;
  }
  return 1;
}

/* worker: `up` = pipe to the parent, `down` = pipe from the parent */
static int worker(int rank, int up, int down)
{
  const int w = 640, h = 480;
  uint8_t *img = malloc((size_t)w * h);
  blobs(img, w, h, 4 * rank); /* the same scene shifted by 4 px per rank */
  if (vksift_loadVulkan() != VKSIFT_SUCCESS)
    return 1;
  vksift_setLogLevel(VKSIFT_LOG_WARNING);
  vksift_Config config = vksift_getDefaultConfig();
  config.input_image_max_size = (uint32_t)(w * h);
  config.max_nb_sift_per_buffer = SLOT_ROWS;
  config.gpu_device_index = 0;
  vksift_Instance inst = NULL;
  if (vksift_createInstance(&inst, &config) != VKSIFT_SUCCESS)
    return 1;
  vksift_detectFeatures(inst, img, w, h, 0u);
  const uint32_t n_own = vksift_getFeaturesNumber(inst, 0u);

  /* set-up: allocate the receive region, hand the handle to the parent, get everybody's back, map the peers */
  uint8_t handle[HANDLE_BYTES], all[NPROC * HANDLE_BYTES];
  if (!vksiftx_exchangeCreate(inst, (uint32_t)rank, NPROC, SLOT_ROWS, handle))
    return 1;
  if (write(up, handle, HANDLE_BYTES) != HANDLE_BYTES || !read_all(down, all, sizeof(all)))
    return 1;
  if (!vksiftx_exchangeConnect(inst, all))
    return 1;
  char token = 1; /* barrier: every region is mapped everywhere before the first push */
  if (write(up, &token, 1) != 1 || !read_all(down, &token, 1))
    return 1;

  /* the step: push + wait + searches against every received block in place, then one download */
  uint32_t counts[NPROC];
  if (!vksiftx_exchangeMatchAllPeers(inst, 0u, counts))
    return 1;
  vksift_Match_2NN *m = malloc(sizeof(vksift_Match_2NN) * (size_t)NPROC * (n_own ? n_own : 1));
  vksiftx_downloadMatchesBlocks(inst, m, NPROC);
  int bad = 0;
  for (int p = 0; p < NPROC; p++)
  {
    if (p == rank)
      continue;
    uint32_t good = 0;
    for (uint32_t i = 0; i < n_own; i++)
    {
      const vksift_Match_2NN *r = &m[(size_t)p * n_own + i];
      if (r->idx_a != i || r->idx_b1 >= counts[p] || r->idx_b2 >= counts[p] || r->idx_b1 == r->idx_b2 || r->dist_a_b1 > r->dist_a_b2)
        bad++;
      if (r->dist_a_b1 < 0.75f * r->dist_a_b2)
        good++;
    }
    printf("rank %d (%u features) vs rank %d (%u features): %u matches passing the ratio test\n", rank, n_own, p, counts[p], good);
    if (good < n_own / 4)
      bad++; /* the images are the same scene shifted by a few pixels */
  }
  fflush(stdout);
  if (write(up, &token, 1) != 1 || !read_all(down, &token, 1)) /* barrier: nobody unmaps a region a peer may still push into */
    return 1;
  vksiftx_exchangeDestroy(inst);
  vksift_destroyInstance(&inst);
  vksift_unloadVulkan();
  free(m);
  free(img);
  return bad ? 2 : 0;
}

int main(void)
{
  int up[NPROC][2], down[NPROC][2];
  pid_t pid[NPROC];
  for (int r = 0; r < NPROC; r++)
  {
    if (pipe(up[r]) != 0 || pipe(down[r]) != 0)
      return 1;
    pid[r] = fork(); /* before any CUDA call: a CUDA context does not survive a fork */
    if (pid[r] == 0)
    {
      close(up[r][0]);
      close(down[r][1]);
      _exit(worker(r, up[r][1], down[r][0]));
    }
    close(up[r][1]);
    close(down[r][0]);
  }
  /* the parent is the "network": gather the handles, broadcast them, then serve two barriers */
  uint8_t all[NPROC * HANDLE_BYTES];
  int ok = 1;
  for (int r = 0; r < NPROC; r++)
    ok &= read_all(up[r][0], all + r * HANDLE_BYTES, HANDLE_BYTES);
  for (int r = 0; r < NPROC && ok; r++)
    ok &= write(down[r][1], all, sizeof(all)) == (ssize_t)sizeof(all);
  for (int b = 0; b < 2 && ok; b++)
  {
    char token;
    for (int r = 0; r < NPROC; r++)
      ok &= read_all(up[r][0], &token, 1);
    for (int r = 0; r < NPROC && ok; r++)
      ok &= write(down[r][1], &token, 1) == 1;
  }
  int failed = !ok;
  for (int r = 0; r < NPROC; r++)
  {
    if (!ok)
      close(down[r][1]); /* unblock workers waiting for us */
    int st = 0;
    waitpid(pid[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0)
      failed = 1;
  }
  printf("exchange_pairs: %s\n", failed ? "FAILED" : "ok");
  return failed;
}
