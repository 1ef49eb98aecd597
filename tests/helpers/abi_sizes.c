/* prints the sizes/offsets the Python binding must agree with (tests/test_abi.py) */
#include <stddef.h>
#include <stdio.h>
#include "vp_b200.h"
int main(void)
{
	printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(vp_camera_model), sizeof(vp_match), sizeof(vp_params), offsetof(vp_params, model),
	       offsetof(vp_params, max_robot_height), offsetof(vp_params, sample_mode), offsetof(vp_match, circ));
	return 0;
}
