"""`python main.py [flags]` -- same entry point and flags as the reference's main.py."""
from irm_motion_planning_b200.main import main, parse_args  # noqa: F401

if __name__ == "__main__":
    main()
