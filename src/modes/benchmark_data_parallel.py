from vdpp_b200.modes.benchmark_data_parallel import *  # noqa: F401,F403
from vdpp_b200.modes.benchmark_data_parallel import main  # noqa: F401

if __name__ == "__main__":
    main()
