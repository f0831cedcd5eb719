"""CPU oracle for the rollout hot path — TEST INFRASTRUCTURE, never imported by the product package."""
