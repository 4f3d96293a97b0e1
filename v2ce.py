"""Entry point with the reference's name: ``python v2ce.py -i clip.mp4 -b 4 ...`` (same flags as
ucsd-hdsi-dvs/V2CE-Toolbox v2ce.py:283-302).  The implementation lives in v2ce_toolbox_b200/v2ce.py."""
from v2ce_toolbox_b200.v2ce import *  # noqa: F401,F403
from v2ce_toolbox_b200.v2ce import main

if __name__ == '__main__':
    main()
