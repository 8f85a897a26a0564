/* C++ caller that turns the error callback into exceptions: the usage the reference documents in
 * src/examples/test_sift_error_handling.cpp:6-16,49-70 (the callback throws, the exception unwinds through the
 * library's C entry point back into the caller's try block; the instance stays usable after an invalid-input error).
 *   g++ -std=c++17 examples/error_handling.cpp -Iinclude -Lvulkansift_b200/lib -lvulkansift -Wl,-rpath,$PWD/vulkansift_b200/lib -o error_handling
 * Prints one line per check and returns 0 when every check behaved as documented. */
#include <vulkansift/vulkansift.h>

#include <cstdio>
#include <stdexcept>
#include <vector>

static void throwing_callback(vksift_Result result)
{
  if (result == VKSIFT_VULKAN_ERROR)
    throw std::runtime_error("device failure inside a vksift function: the instance must be destroyed");
  if (result == VKSIFT_INVALID_INPUT_ERROR)
    throw std::invalid_argument("invalid argument given to a vksift function: the instance can still be used");
}

int main()
{
  const uint32_t NB_BUFF = 5;
  vksift_setLogLevel(VKSIFT_LOG_WARNING);
  if (vksift_loadVulkan() != VKSIFT_SUCCESS)
    return 2;
  vksift_Config config = vksift_getDefaultConfig();
  config.on_error_callback_function = throwing_callback;
  config.sift_buffer_count = NB_BUFF;
  config.input_image_max_size = 640 * 480;
  vksift_Instance inst = NULL;
  if (vksift_createInstance(&inst, &config) != VKSIFT_SUCCESS)
    return 2;
  int failures = 0;

  /* any buffer index >= NB_BUFF must end in the caller's catch block, every index below must not throw */
  uint32_t first_bad = 0xffffffffu;
  try
  {
    for (uint32_t i = 0; i < NB_BUFF * 2; i++)
    {
      first_bad = i;
      vksift_getFeaturesNumber(inst, i);
    }
    first_bad = 0xffffffffu;
  }
  catch (std::invalid_argument &e)
  {
    std::printf("buffer %u: std::invalid_argument caught: %s\n", first_bad, e.what());
  }
  if (first_bad != NB_BUFF)
  {
    std::printf("FAIL: expected the first exception at buffer %u, got %u\n", NB_BUFF, first_bad);
    failures++;
  }

  /* an image larger than input_image_max_size: same path through vksift_detectFeatures */
  std::vector<uint8_t> big(1000 * 1000, 0);
  bool thrown = false;
  try
  {
    vksift_detectFeatures(inst, big.data(), 1000, 1000, 0);
  }
  catch (std::invalid_argument &)
  {
    thrown = true;
  }
  std::printf("oversized image: %s\n", thrown ? "std::invalid_argument caught" : "NO EXCEPTION");
  failures += thrown ? 0 : 1;

  /* the instance is still usable after invalid-input errors */
  std::vector<uint8_t> img(640 * 480);
  for (size_t i = 0; i < img.size(); i++)
    img[i] = (uint8_t)(((i % 640) / 16 + (i / 640) / 16) % 2 ? 200 : 40);
  uint32_t n = 0;
  try
  {
    vksift_detectFeatures(inst, img.data(), 640, 480, 1);
    n = vksift_getFeaturesNumber(inst, 1);
  }
  catch (std::exception &e)
  {
    std::printf("FAIL: valid call threw: %s\n", e.what());
    failures++;
  }
  std::printf("valid detection after the errors: %u features\n", n);
  failures += n > 0 ? 0 : 1;

  vksift_destroyInstance(&inst);
  vksift_unloadVulkan();
  std::printf(failures ? "error_handling: FAILED\n" : "error_handling: ok\n");
  return failures ? 1 : 0;
}
