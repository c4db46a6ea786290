"""oracle -- TEST INFRASTRUCTURE (CPU checker). Never imported by onesolver_b200/."""
