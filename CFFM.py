"""Drop-in entry point: ``python CFFM.py --dataset frappe ...`` with the reference's flags."""
from cffm_b200.cli import main

if __name__ == "__main__":
    main()
