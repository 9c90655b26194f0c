python scripts/train_glue_profile.py 2>&1 | tail -44
