import os
exec(os.environ["H264B2_TEST_CODE"])
