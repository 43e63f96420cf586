"""CPU oracle for redback_b200 -- TEST INFRASTRUCTURE ONLY (never imported by the product)."""
