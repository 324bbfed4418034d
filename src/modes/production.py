from vdpp_b200.modes.production import *  # noqa: F401,F403
from vdpp_b200.modes.production import main  # noqa: F401

if __name__ == "__main__":
    main()
