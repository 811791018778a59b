"""`python -m training.main ...` with the reference's flags (scripts/*.sh) -> clipself_b200.training.main."""
from clipself_b200.training.main import main

if __name__ == "__main__":
    main()
